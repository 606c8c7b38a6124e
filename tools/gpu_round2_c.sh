#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/ub_intops.bin > gpurun_out/r02_ub_intops.jsonl 2>&1
grep -E "IMAD|residue|IDP.4A" gpurun_out/r02_ub_intops.jsonl
echo "--- pred default (pairs)"; timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- pred 14 moduli"; GPE_OZAKI=14 timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- llh default"; timeout 300 python tools/perf_llh.py 4096 16 32 5 2>&1 | tail -n 3
echo "--- llh cluster_i=2"; GPE_OZAKI_CLUSTER_I=2 timeout 300 python tools/perf_llh.py 4096 16 32 5 2>&1 | tail -n 3
echo "--- llh 14 moduli"; GPE_OZAKI=14 timeout 300 python tools/perf_llh.py 4096 16 32 5 2>&1 | tail -n 3
echo "--- llh default again"; timeout 300 python tools/perf_llh.py 4096 16 32 5 2>&1 | tail -n 3
