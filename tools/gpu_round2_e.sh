#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_predict.py tests/test_gpu_hm.py tests/test_gpu_ozaki.py -m gpu -x -q 2>&1 | tail -n 12
echo "--- pred fixed scale (default)"; timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
echo "--- pred searched maxima"; GPE_PRED_FIXED_SCALE=0 timeout 300 python tools/perf_pred.py 2000 8 4194304 2>&1 | tail -n 2
