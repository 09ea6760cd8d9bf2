#!/bin/bash
for defs in "PREP_SMALL_MULT=2" "PREP_SMALL_MULT=3" "PREP_SMALL_MULT=4"; do
  echo "== $defs"
  M2T_DEFS="$defs" timeout 900 python -m m2trans_b200.build --force > /dev/null 2>&1 || echo build failed
  for c in cfg2 frame2 cfg1; do timeout 300 python tools/stage_profile.py $c 2>&1 | grep "branch_prep_all\|replayed" | cut -c1-40 | tr '\n' ' '; echo; done
done
