#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_stage_attnz_gpu.py -x -q -k "fused_tail or golden or reproducible" > gpurun_out/r2j_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2j_tests.log; tail -3 gpurun_out/r2j_tests.log
for w in cfg2 cfg1; do timeout 300 python bench.py --no-cpu --workload $w > gpurun_out/r2j_$w.log 2>&1; echo "$w: $(tail -1 gpurun_out/r2j_$w.log | python tools/show_bench.py)"; done
timeout 300 python tools/stage_profile.py cfg2 2>&1 | cut -c1-120 | grep -E "tail|total"
SAN_TOOL=memcheck bash tools/gpu_sanitize.sh > /dev/null 2>&1; tail -4 gpurun_out/sanitize_memcheck.log
SAN_TOOL=synccheck bash tools/gpu_sanitize.sh > /dev/null 2>&1; tail -3 gpurun_out/sanitize_synccheck.log
