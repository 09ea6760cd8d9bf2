#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_stage_attnz_gpu.py tests/test_graph_gpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2n_tests.log
for c in cfg1 frame2 frame4; do timeout 300 python tools/stage_profile.py $c 2>&1 | cut -c1-130 | tail -9 | tee gpurun_out/r2n_stage_$c.log; M2T_AZ_PAIRED=1 timeout 300 python tools/stage_profile.py $c 2>&1 | cut -c1-130 | tail -9 | tee gpurun_out/r2n_stage_${c}_paired.log; done
timeout 300 python bench.py --workload cfg1 --no-cpu 2>&1 | tail -1 > gpurun_out/r2n_cfg1.log; python tools/show_bench.py < gpurun_out/r2n_cfg1.log
