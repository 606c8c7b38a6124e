#!/bin/bash
# A/B of the two-CTAs-per-SM LAUUM + gradient kernel (GPE_LG64=1) on one box: parity tests, then the n = 4096 step.
GPE_LG64=1 python -m pytest tests/test_gpu_core.py tests/test_gpu_headline_golden.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
for v in 0 1; do
  echo "LG64=$v"
  GPE_LG64=$v python tools/perf_llh.py 4096 16 32 5 | tail -2 | head -1
  GPE_LG64=$v GPE_STREAMS=1 GPE_GRAPHS=0 python tools/perf_llh.py 4096 16 32 3 | tail -2 | head -1
  GPE_LG64=$v python tools/perf_llh.py 1000 8 64 5 | tail -2 | head -1
done
