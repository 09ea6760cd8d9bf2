#!/bin/bash
# two-GPU checks: nn.DataParallel replicas + both bench arms under torchrun
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --no-header -k "multi_gpu" 2>&1 | tail -2
bash tools/gpu_two.sh
