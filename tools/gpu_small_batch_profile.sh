#!/bin/bash
# launch list of a single-item n=4096 evaluation (the ragged tail of a multistart): where do 5 ms go?
GPE_GRAPHS=0 python tools/perf_llh.py 4096 16 1 3 > gpurun_out/r02_plain_b1.log 2>&1 &&
GPE_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b1.csv python tools/perf_llh.py 4096 16 1 3 > gpurun_out/r02_ncu_b1.log 2>&1
tail -n 4 gpurun_out/r02_plain_b1.log
