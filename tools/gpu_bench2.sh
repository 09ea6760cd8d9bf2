#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
echo "== graph"; python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'], d['e2e'])"
echo "== eager"; M2T_CUDA_GRAPH=0 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'], d['e2e'])"
echo "== eager nopdl"; M2T_NO_PDL=1 M2T_CUDA_GRAPH=0 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'], d['e2e'])"
echo "== graph nopdl"; M2T_NO_PDL=1 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'], d['e2e'])"
