#!/bin/bash
for sw in "3 3" "6 3" "20 5"; do set -- $sw; python bench.py --steps $1 --warmup $2 --no-extra --no-cpu --grid-points 3e5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps',d['steps'],'warmup',d['warmup'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1))"; done
