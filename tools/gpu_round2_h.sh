#!/bin/bash
bash tools/gpu_profile_pred_int8.sh
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2_smoke.log
timeout 1500 python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_final2.err
