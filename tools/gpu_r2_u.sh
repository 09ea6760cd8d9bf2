#!/bin/bash
# sanitizers over the final kernels (memcheck + synccheck), then the whole -m gpu suite
mkdir -p gpurun_out
SAN_TOOL=memcheck SAN_TAIL=8 bash tools/gpu_sanitize.sh
SAN_TOOL=synccheck SAN_TAIL=8 bash tools/gpu_sanitize.sh
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2u_tests.log
