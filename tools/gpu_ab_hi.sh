#!/bin/bash
# the residue GEMM on a high-priority twin stream (GPE_OZAKI_HI): eager launches and graph replay
llh() { python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --grid-points 3e5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('evals/s',round(d['value'],1),'ms/step',round(d['ms_per_step'],2))"; }
echo "--- eager, hi"; GPE_GRAPHS=0 llh
echo "--- eager, no hi"; GPE_GRAPHS=0 GPE_OZAKI_HI=0 llh
echo "--- eager, hi, 4 groups"; GPE_GRAPHS=0 GPE_OZAKI_STREAMS=4 llh
echo "--- eager, no hi, 4 groups"; GPE_GRAPHS=0 GPE_OZAKI_HI=0 GPE_OZAKI_STREAMS=4 llh
echo "--- overlap tool, hi"; timeout 300 python tools/oz_overlap.py 2>&1 | tail -n 2 | head -n 1
echo "--- overlap tool, no hi"; GPE_OZAKI_HI=0 timeout 300 python tools/oz_overlap.py 2>&1 | tail -n 2 | head -n 1
