#!/bin/bash
# Round-2 ncu evidence (run on the GPU box from the repo root; one gpurun call):
#  (1) launch list of the bench command with serial launches (per-launch device time of every kernel of a step),
#  (2) --set full capture of the dominant kernel (LAUUM + gradient epilogue, one 32-item launch),
#  (3) --set full capture of the prediction chunk's kernels (tabulated cross-covariance, skinny panel, TRMM with column norms).
set -x
BENCH="python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu --no-extra --grid-points 3e5"
$BENCH > gpurun_out/r02_plain_bench.json 2> gpurun_out/r02_plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 420 --csv --log-file gpurun_out/r02_launches_bench_serial.csv $BENCH > gpurun_out/r02_ncu_bench.log 2>&1
GPE_STREAMS=1 GPE_GRAPHS=0 python tools/perf_llh.py 4096 16 32 1 > gpurun_out/r02_plain_llh.log 2>&1 &&
GPE_STREAMS=1 GPE_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:lauum_grad -s 1 -c 1 -o gpurun_out/r02_ncu_lauum_grad_b32 python tools/perf_llh.py 4096 16 32 1 > gpurun_out/r02_ncu_llh.log 2>&1
python tools/perf_pred.py 2000 8 131072 > gpurun_out/r02_plain_pred.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"xcov_grid|gemm_dmma" -s 6 -c 3 -o gpurun_out/r02_ncu_predict_chunk python tools/perf_pred.py 2000 8 131072 > gpurun_out/r02_ncu_pred.log 2>&1
python -m pytest tests/test_gpu_headline_golden.py -m gpu -q -s 2>&1 | grep -i "worst\|passed\|failed" > gpurun_out/r02_headline_worst.txt
tail -3 gpurun_out/r02_plain_llh.log gpurun_out/r02_plain_pred.log gpurun_out/r02_ncu_llh.log gpurun_out/r02_ncu_pred.log gpurun_out/r02_headline_worst.txt
