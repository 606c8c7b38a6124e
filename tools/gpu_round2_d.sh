#!/bin/bash
for s in 1 2 3 4 8; do echo "--- llh oz streams=$s"; GPE_OZAKI_STREAMS=$s timeout 300 python tools/perf_llh.py 4096 16 32 6 2>&1 | tail -n 3 | head -n 2; done
echo "--- overlap tool"; timeout 300 python tools/oz_overlap.py 2>&1 | tail -n 3
