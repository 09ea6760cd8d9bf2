#!/bin/bash
# parity of the default path on uniform and speckle inputs over seeds (tests/parity_sweep.py) + bench
mkdir -p gpurun_out
timeout 600 python tests/parity_sweep.py 3 200 266 3 2>&1 | tail -7 | tee gpurun_out/parity_x3.log
timeout 300 python tests/parity_sweep.py 3 64 64 3 2>&1 | tail -1
timeout 400 python tests/parity_sweep.py 2 128 160 3 2>&1 | tail -1
timeout 400 python tests/parity_sweep.py 4 128 128 3 2>&1 | tail -1
for i in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python tools/show_bench.py; done
