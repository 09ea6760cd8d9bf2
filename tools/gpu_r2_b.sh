#!/bin/bash
# image groups: correctness (graph / batch-consistency tests) and cfg2 timing for 1, 2, 4 groups
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_engine_gpu.py -q -x -k "graph or batch_consistency or oracle or full_size_cfg2" > gpurun_out/r2b_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_tests.log
tail -4 gpurun_out/r2b_tests.log
for g in 1 2 4 8; do
  M2T_IMAGE_GROUPS=$g timeout 300 python bench.py --no-cpu > gpurun_out/r2b_bench_g$g.log 2>&1
  echo "groups $g: $(tail -1 gpurun_out/r2b_bench_g$g.log | python tools/show_bench.py)"
done
M2T_IMAGE_GROUPS=1 timeout 300 python bench.py --no-cpu --workload cfg1 > gpurun_out/r2b_cfg1.log 2>&1; echo "cfg1: $(tail -1 gpurun_out/r2b_cfg1.log | python tools/show_bench.py)"
for g in 1 2; do
M2T_IMAGE_GROUPS=$g timeout 600 python bench.py --no-cpu --workload cfg3 > gpurun_out/r2b_cfg3_g$g.log 2>&1; echo "cfg3 g$g: $(tail -1 gpurun_out/r2b_cfg3_g$g.log | python tools/show_bench.py)"
done
M2T_IMAGE_GROUPS=1 timeout 600 python bench.py --no-cpu --workload cfg4 > gpurun_out/r2b_cfg4.log 2>&1; echo "cfg4: $(tail -1 gpurun_out/r2b_cfg4.log | python tools/show_bench.py)"
