#!/bin/bash
# tuning: tail_strip's idle-role sleep quanta (rebuild on the box), tail time from the eager stage profile + replayed forward
mkdir -p gpurun_out
for defs in "TS_MMA_NS=200 TS_IDLE_NS=200" "TS_MMA_NS=32 TS_IDLE_NS=200" "TS_MMA_NS=0 TS_IDLE_NS=200" "TS_MMA_NS=32 TS_IDLE_NS=64" "TS_MMA_NS=0 TS_IDLE_NS=0"; do
  echo "== $defs"
  M2T_DEFS="$defs" timeout 900 python -m m2trans_b200.build --force > /dev/null 2>&1 || echo build failed
  timeout 300 python tools/stage_profile.py cfg2 2>&1 | grep "tail_strip\|replayed" | cut -c1-40
  timeout 300 python tools/stage_profile.py cfg2 2>&1 | grep "tail_strip\|replayed" | cut -c1-40
done 2>&1 | tee gpurun_out/r2t_sweep.log
