#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_tests.log
tail -5 gpurun_out/r2e_tests.log
timeout 300 python bench.py --no-cpu > gpurun_out/r2e_bench.log 2>&1; echo "cfg2: $(tail -1 gpurun_out/r2e_bench.log | python tools/show_bench.py)"
timeout 300 python tools/stage_profile.py cfg2 2>&1 | cut -c1-100 > gpurun_out/r2e_stage_cfg2.log; cat gpurun_out/r2e_stage_cfg2.log
for w in cfg1 cfg3 cfg4; do timeout 600 python bench.py --no-cpu --workload $w > gpurun_out/r2e_$w.log 2>&1; echo "$w: $(tail -1 gpurun_out/r2e_$w.log | python tools/show_bench.py)"; done
