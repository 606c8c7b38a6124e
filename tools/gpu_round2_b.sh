#!/bin/bash
# same-box checks: instruction rates, knob A/Bs of the INT8 route, bench line
mkdir -p gpurun_out


echo "--- pred cluster_i=2"; GPE_OZAKI_CLUSTER_I=2 timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- pred default"; timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- pred 14 moduli"; GPE_OZAKI=14 timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- llh default"; timeout 300 python tools/perf_llh.py 4096 16 32 4 2>&1 | tail -n 3
echo "--- llh oz_min=512"; GPE_OZAKI_MIN=512 timeout 300 python tools/perf_llh.py 4096 16 32 4 2>&1 | tail -n 3
echo "--- llh 14 moduli"; GPE_OZAKI=14 timeout 300 python tools/perf_llh.py 4096 16 32 4 2>&1 | tail -n 3
timeout 1500 python bench.py > gpurun_out/r2_bench_predint8.json 2> gpurun_out/r2_bench_predint8.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_predint8.err
