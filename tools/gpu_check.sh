#!/bin/bash
# One gpurun call: probes, parity tests, smoke, short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
echo "== probes (one process per probe: a faulting hypothesis must not poison the others)"
: > gpurun_out/probes.log
for t in $(grep -o "^def test_[a-z0-9_]*" tests/test_probes.py | sed "s/def //"); do
  timeout 120 python -m pytest "tests/test_probes.py::$t" -m gpu -q -rA --no-header 2>&1 | grep -E "^\[probe\]|passed|failed|Error" | tee -a gpurun_out/probes.log
done
echo "== engine"; timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -rA --no-header 2>&1 | tail -80 | tee gpurun_out/engine.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
