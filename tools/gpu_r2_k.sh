#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2k_tests.log; tail -4 gpurun_out/r2k_tests.log
timeout 300 python bench.py --no-cpu > gpurun_out/r2k_bench.log 2>&1; echo "cfg2: $(tail -1 gpurun_out/r2k_bench.log | python tools/show_bench.py)"
for t in memcheck synccheck racecheck; do SAN_TOOL=$t SAN_TAIL=400 bash tools/gpu_sanitize.sh > /dev/null 2>&1; echo "== $t"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_$t.log | sort | uniq -c | sort -rn | head -6; done
