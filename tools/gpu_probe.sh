#!/bin/bash
mkdir -p gpurun_out
for t in "$@"; do
  timeout 120 python -m pytest "tests/test_probes.py::$t" -m gpu -q -rA --no-header 2>&1 | grep -E "^\[probe\]|passed|failed|Error" | tee -a gpurun_out/probes2.log
done
