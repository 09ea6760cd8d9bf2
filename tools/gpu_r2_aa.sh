#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_graph_gpu.py tests/test_stage_attnz_gpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2aa_tests.log
for c in cfg2 cfg3 cfg1; do timeout 300 python tools/stage_profile.py $c 2>&1 | grep "attn_z_kernelILi64\|replayed\|total" | cut -c1-100; done | tee gpurun_out/r2aa_stage.log
for i in 1 2; do timeout 300 python bench.py --no-cpu 2>&1 | tail -1 > gpurun_out/r2aa_cfg2_$i.log; python tools/show_bench.py < gpurun_out/r2aa_cfg2_$i.log; done
for c in cfg4 cfg3 cfg1; do timeout 300 python bench.py --workload $c --no-cpu 2>&1 | tail -1 > gpurun_out/r2aa_$c.log; python tools/show_bench.py < gpurun_out/r2aa_$c.log; done
