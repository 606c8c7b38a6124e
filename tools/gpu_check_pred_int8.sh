#!/bin/bash
# INT8 route of the prediction product: full GPU suite, then the launch list of two prediction chunks (run from the repo root).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pred_int8_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pred_int8_tests.log
tail -n 15 gpurun_out/r2_pred_int8_tests.log
CMD="python tools/perf_pred.py 2000 8 131072"
timeout 300 $CMD > gpurun_out/r02_plain_pred_int8.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_pred_int8.csv $CMD > gpurun_out/r02_ncu_pred_int8_list.log 2>&1
tail -n 3 gpurun_out/r02_plain_pred_int8.log
