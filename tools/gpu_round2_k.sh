#!/bin/bash
run() { python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --grid-points 3e5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['by_kernel_ms_per_step']; print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'grad',round(r['grad_reduce'],2),'cov',round(r['cov_build'],2))"; }
echo "--- E copy (default)"; run
echo "--- recomputed exp"; GPE_GRAD_E=0 run
echo "--- E copy (default)"; run
timeout 900 python -m pytest tests/test_gpu_headline_golden.py tests/test_gpu_core.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -n 3
