#!/bin/bash
# tuning: rebuild the library with different Linear-kernel ring depths on the GPU box and time tools/lin_timing.py
for defs in "LG_STAGES_OVERRIDE=3" "LG_STAGES_OVERRIDE=4"; do
  echo "== $defs"
  M2T_DEFS="$defs" timeout 600 python -m m2trans_b200.build --force > /dev/null 2>&1 || echo build failed
  timeout 200 python tools/lin_timing.py 2>&1 | grep "^M " | head -6 | cut -c1-200
done
