#!/bin/bash
# ncu evidence for one prediction chunk on the INT8 route (run on the GPU box from the repo root; one gpurun call):
#  (0) the full GPU suite on this build,
#  (1) launch list of tools/perf_pred.py (two 65 536-point chunks per call, three calls),
#  (2) --set full capture of the seven launches of the third call's first chunk, selected by launch index (the launch order is
#      deterministic: grid points, cross-covariance, skinny panel, residue conversion of the slab, residue GEMM, CRT with column
#      norms, finalize -- the residue planes of L^-1 were made by the first call).
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_final_tests.log
tail -n 4 gpurun_out/r2_final_tests.log
CMD="python tools/perf_pred.py 2000 8 131072"
timeout 300 $CMD > gpurun_out/r02_plain_pred_int8.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_pred_int8.csv $CMD > gpurun_out/r02_ncu_pred_int8_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 126 --launch-count 7 \
    -o gpurun_out/r02_ncu_predict_chunk_int8 -f $CMD > gpurun_out/r02_ncu_pred_int8.log 2>&1
tail -n 3 gpurun_out/r02_plain_pred_int8.log gpurun_out/r02_ncu_pred_int8.log
