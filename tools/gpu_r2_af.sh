#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_graph_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2af_tests.log
for c in cfg2 cfg3 cfg4; do timeout 300 python tools/stage_profile.py $c 2>&1 | grep "tail_up\|replayed" | cut -c1-60 | tr '\n' ' '; echo; done
for i in 1 2; do timeout 300 python bench.py --no-cpu 2>&1 | tail -1 | python tools/show_bench.py; done
