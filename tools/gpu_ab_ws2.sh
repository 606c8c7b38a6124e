#!/bin/bash
# A/B of the two-CTAs-per-SM GEMM variant (GPE_WS2=1) on one box: correctness (GPU tests) and the two headline paths.
GPE_WS2=1 python -m pytest tests/test_gpu_core.py tests/test_gpu_predict.py tests/test_gpu_headline_golden.py -m gpu -q -x 2>&1 | tail -3
for v in 0 1; do
  echo "WS2=$v"
  GPE_WS2=$v python tools/perf_llh.py 4096 16 32 5 | tail -2 | head -1
  GPE_WS2=$v python tools/perf_llh.py 1000 8 64 5 | tail -2 | head -1
  GPE_WS2=$v python tools/perf_pred.py 2000 8 4194304 | tail -1
done
