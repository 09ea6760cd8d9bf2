#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2ag_tests.log
for i in 1 2; do timeout 300 python bench.py --no-cpu 2>&1 | tail -1 | python tools/show_bench.py; done
for c in cfg1 cfg3 cfg4; do timeout 300 python bench.py --workload $c --no-cpu 2>&1 | tail -1 | python tools/show_bench.py; done
timeout 300 python tools/bench_clip.py --batch 256 2>&1 | tail -1 | cut -c1-260
timeout 300 python tools/bench_clip.py --batch 32 2>&1 | tail -1 | cut -c1-260
