#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_predict.py tests/test_gpu_hm.py -m gpu -x -q 2>&1 | tail -n 12
echo "--- pred swapped roles (default)"; timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- pred unswapped"; GPE_OZAKI_SUMSQ_SWAP=0 timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- pred swapped, pairs"; GPE_OZAKI_CLUSTER=2 timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- pred swapped, chunk 131072"; GPE_PRED_CHUNK=131072 timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- pred swapped, chunk 32768"; GPE_PRED_CHUNK=32768 timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
