#!/bin/bash
for rep in 1 2; do
for cl in 4 2; do echo "--- llh cluster=$cl (rep $rep)"; GPE_OZAKI_CLUSTER=$cl timeout 300 python tools/perf_llh.py 4096 16 32 12 2>&1 | grep iter | tail -n 8 | awk '{print $3}' | sort -n | tr '\n' ' '; echo; done
done
echo "--- llh cluster=4, row-triangular 2"; GPE_OZAKI_CLUSTER_I=2 timeout 300 python tools/perf_llh.py 4096 16 32 12 2>&1 | grep iter | tail -n 8 | awk '{print $3}' | sort -n | tr '\n' ' '; echo
echo "--- pred cluster=2"; GPE_OZAKI_CLUSTER=2 timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- pred cluster=4"; timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
