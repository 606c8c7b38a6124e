#!/bin/bash
# Eight-GPU check of the bench command (run on the GPU box from the repo root).
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_g8.json 2> gpurun_out/r2_bench_g8.err
tail -c 400 gpurun_out/r2_bench_g8.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2_bench_g8.json"))
print(j["n_gpus"], j["value"], j["ms_per_step"], j["roofline"]["frac"])
e = j["extra"]
print(e.get("posterior_preds_per_s"), e["posterior_roofline"]["frac"] if "posterior_roofline" in e else None)
print(e.get("history_match")); print(e.get("config3_optimisation")); print(e.get("extra_configs_error"), e.get("posterior_error"))
PY
