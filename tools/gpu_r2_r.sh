#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_graph_gpu.py tests/test_loader_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2r_tests.log
timeout 300 python tools/stage_profile.py cfg2 2>&1 | cut -c1-110 | tail -10 | tee gpurun_out/r2r_stage_cfg2.log
for i in 1 2; do timeout 300 python bench.py --no-cpu 2>&1 | tail -1 > gpurun_out/r2r_cfg2_$i.log; python tools/show_bench.py < gpurun_out/r2r_cfg2_$i.log; done
timeout 300 python bench.py --workload cfg4 --no-cpu 2>&1 | tail -1 > gpurun_out/r2r_cfg4.log; python tools/show_bench.py < gpurun_out/r2r_cfg4.log
timeout 300 python bench.py --workload cfg3 --no-cpu 2>&1 | tail -1 > gpurun_out/r2r_cfg3.log; python tools/show_bench.py < gpurun_out/r2r_cfg3.log
