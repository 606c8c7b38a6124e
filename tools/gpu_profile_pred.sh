#!/bin/bash
# ncu evidence for one prediction chunk (run on the GPU box from the repo root; one gpurun call):
#  (1) launch list of tools/perf_pred.py (two 65 536-point chunks per call, three calls),
#  (2) --set full capture of one chunk's kernels, selected by launch index (the small GEMMs of gpe_fit_state share the
#      function names of the chunk's GEMMs, so a name filter picks the wrong launches).
set -x
CMD="python tools/perf_pred.py 2000 8 131072"
$CMD > gpurun_out/r02_plain_pred.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_pred.csv $CMD > gpurun_out/r02_ncu_pred_list.log 2>&1
# the launch order is deterministic (116 launches; tools/launch_summary.py on the list above): launches 94..99 are the
# second call's first chunk -- factor tables, grid points, cross-covariance, skinny panel, TRMM with column norms, finalize.
# (Selecting by template arguments does not work: ncu's -k matches another spelling of them than the one it prints.)
ncu --set full --clock-control none --import-source on --launch-skip 94 --launch-count 6 \
    -o gpurun_out/r02_ncu_predict_chunk -f $CMD > gpurun_out/r02_ncu_pred.log 2>&1
tail -n 4 gpurun_out/r02_plain_pred.log gpurun_out/r02_ncu_pred.log
