#!/bin/bash
mkdir -p gpurun_out
for defs in "PREP_GRID_MULT=8" "PREP_GRID_MULT=16" "PREP_GRID_MULT=4"; do
  echo "== $defs"
  M2T_DEFS="$defs" timeout 900 python -m m2trans_b200.build --force > /dev/null 2>&1 || echo build failed
  for c in cfg2 cfg3 cfg4; do timeout 300 python tools/stage_profile.py $c 2>&1 | grep "branch_prep_all\|replayed" | cut -c1-40 | tr '\n' ' '; echo; done
done 2>&1 | tee gpurun_out/r2ad_prep.log
