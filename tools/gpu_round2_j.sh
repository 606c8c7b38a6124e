#!/bin/bash
run() { python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --grid-points 3e5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2))"; }
echo "--- default"; run
echo "--- asym priorities"; GPE_SUB_ASYM=1 run
echo "--- asym priorities, 4 groups"; GPE_SUB_ASYM=1 GPE_OZAKI_STREAMS=4 run
echo "--- 4 groups"; GPE_OZAKI_STREAMS=4 run
echo "--- default"; run
