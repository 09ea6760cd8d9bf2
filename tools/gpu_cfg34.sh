#!/bin/bash
for w in cfg4 cfg3 cfg4; do timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python tools/show_bench.py; done
timeout 300 python tools/stage_profile.py cfg4 | c++filt | cut -c1-100
