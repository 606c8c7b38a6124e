#!/bin/bash
# ncu evidence for one prediction chunk (run on the GPU box from the repo root; one gpurun call):
#  (1) launch list of tools/perf_pred.py (two 65 536-point chunks per call, three calls),
#  (2) --set full capture of the chunk's kernels, selected by their full template names so that the small GEMMs of
#      gpe_fit_state (same function names, other template arguments) are not picked up.
set -x
CMD="python tools/perf_pred.py 2000 8 131072"
$CMD > gpurun_out/r02_plain_pred.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_pred.csv $CMD > gpurun_out/r02_ncu_pred_list.log 2>&1
SEL='regex:xcov_grid_kernel|grid_table_kernel|gemm_dmma_ws_kernel<1, 0, 1|gemm_dmma_kernel<16, 128|predict_finalize_kernel'
# skip the first call's chunks (cold), take one whole chunk of the second call
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$SEL" -s 10 -c 5 \
    -o gpurun_out/r02_ncu_predict_chunk $CMD > gpurun_out/r02_ncu_pred.log 2>&1
tail -4 gpurun_out/r02_plain_pred.log gpurun_out/r02_ncu_pred.log
grep -c gemm_dmma_ws_kernel gpurun_out/r02_launches_pred.csv
