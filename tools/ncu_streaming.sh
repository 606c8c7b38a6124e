#!/bin/bash
# Durations and pipe utilisation of the streaming kernels (covariance build, gradient reduction, cross-covariance)
# for one likelihood batch of 4 items at n = 4096, d = 16 and one prediction chunk at n = 2000, d = 8.
# Writes gpurun_out/streaming_{llh,pred}.csv (ncu --csv); run under gpurun, one GPU.
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
GPE_GRAPHS=0 GPE_STREAMS=1 ncu --metrics $M --clock-control none -k regex:"grad_partial|cov_build" -c 2 --csv --log-file gpurun_out/streaming_llh.csv python tools/perf_llh.py 4096 16 4 0 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:"xcov" -c 1 --csv --log-file gpurun_out/streaming_pred.csv python tools/perf_pred.py 2000 8 65536 > /dev/null 2>&1
