#!/bin/bash
# ncu evidence for the INT8 tensor-core route (run on the GPU box from the repo root; one gpurun call):
#  (1) launch list of the bench command with serial launches (per-launch device time of every kernel of a step),
#  (2) --set full capture of the dominant kernel, oz_gemm_kernel: the 13 launches of one step (top level, second level, LAUUM last),
#  (3) --set full of one launch each of the conversion and CRT kernels.
set -x
BENCH="python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu --no-extra --grid-points 3e5"
$BENCH > gpurun_out/r02_plain_bench_int8.json 2> gpurun_out/r02_plain_bench_int8.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 420 --csv --log-file gpurun_out/r02_launches_bench_int8_serial.csv $BENCH > gpurun_out/r02_ncu_bench_int8.log 2>&1
export GPE_STREAMS=1 GPE_GRAPHS=0
python tools/perf_llh.py 4096 16 32 1 > gpurun_out/r02_plain_llh_int8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -s 13 -c 13 -o gpurun_out/r02_ncu_oz_gemm_b32 -f python tools/perf_llh.py 4096 16 32 1 > gpurun_out/r02_ncu_llh_int8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"oz_convert|oz_combine" -s 52 -c 6 -o gpurun_out/r02_ncu_oz_prepost_b32 -f python tools/perf_llh.py 4096 16 32 1 > gpurun_out/r02_ncu_llh_int8b.log 2>&1
tail -n 2 gpurun_out/r02_plain_llh_int8.log gpurun_out/r02_ncu_llh_int8.log gpurun_out/r02_ncu_llh_int8b.log
