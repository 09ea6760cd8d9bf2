#!/bin/bash
# sanitizers over the staged glue (memcheck, synccheck, racecheck summary), bit-reproducibility stress, then timing stamps
mkdir -p gpurun_out
SAN_TOOL=memcheck SAN_TAIL=8 bash tools/gpu_sanitize.sh
SAN_TOOL=synccheck SAN_TAIL=8 bash tools/gpu_sanitize.sh
timeout 600 python -m pytest tests/test_stage_attnz_gpu.py tests/test_engine_gpu.py -x -q -k "reproducible or determinism or single_window" 2>&1 | tail -3
