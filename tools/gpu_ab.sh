#!/bin/bash
mkdir -p gpurun_out
echo "graph, no PDL"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python tools/show_bench.py
echo "graph + PDL"; M2T_GRAPH_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python tools/show_bench.py
echo "eager + PDL"; M2T_CUDA_GRAPH=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python tools/show_bench.py
