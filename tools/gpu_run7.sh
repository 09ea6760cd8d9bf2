#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -4 | tee gpurun_out/all.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench.log
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.log
