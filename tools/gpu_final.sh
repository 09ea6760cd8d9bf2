#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/final_ref.log
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/final_bench.json | cut -c1-300
