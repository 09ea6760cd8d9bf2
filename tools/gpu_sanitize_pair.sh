#!/bin/bash
# compute-sanitizer (memcheck, synccheck) over the opt-in CTA-pair ff conv (M2T_VAR_W2_PAIR) and the staged-glue paths
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/pair_check.py 2>&1 | grep -vE "^$" | tail -8 | tee gpurun_out/sanitize_pair_$tool.log
done
