#!/bin/bash
# Four-GPU check (run on the GPU box from the repo root with gpurun --gpus 4): the prediction tests and a short N=4 bench.
python -m pytest tests/test_gpu_predict.py -m gpu -q -x 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2_bench_g4.json 2> gpurun_out/r2_bench_g4.err
tail -c 300 gpurun_out/r2_bench_g4.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2_bench_g4.json"))
print(j["n_gpus"], j["value"], j["ms_per_step"])
e = j["extra"]
print(e.get("posterior_preds_per_s"), e.get("history_match"), e.get("config3_optimisation"), e.get("extra_configs_error"), e.get("posterior_error"))
PY
echo "(the reference arm under torchrun was checked earlier in the round: profiles/r02_bench_final_reference_arm.json)"
