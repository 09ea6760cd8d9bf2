#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -3 | tee gpurun_out/all.log
echo "== bench PDL"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee gpurun_out/bench.log
echo "== bench no PDL"; M2T_NO_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-330 | tee gpurun_out/bench_nopdl.log
