#!/bin/bash
mkdir -p gpurun_out
echo "== probes"
: > gpurun_out/probes.log
for t in $(grep -o "^def test_[a-z0-9_]*" tests/test_probes.py | sed "s/def //"); do
  timeout 120 python -m pytest "tests/test_probes.py::$t" -m gpu -q -rA --no-header 2>&1 | grep -E "^\[probe\]|passed|failed|Error" | tee -a gpurun_out/probes.log
done
echo "== qkv umma"; timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -rA --no-header -k "stage_qkv" 2>&1 | tail -30 | tee gpurun_out/qkv.log
echo "== engine"; timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -k "not stage_qkv" 2>&1 | tail -30 | tee gpurun_out/engine.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench.log
