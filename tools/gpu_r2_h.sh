#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_loader_gpu.py tests/test_rlutrans_gpu.py -x -q -k "ffconv or golden or cftm or tblock or loader or images or submodules or eval_loop" > gpurun_out/r2h_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2h_tests.log; tail -4 gpurun_out/r2h_tests.log
timeout 300 python bench.py --no-cpu > gpurun_out/r2h_bench.log 2>&1; echo "cfg2: $(tail -1 gpurun_out/r2h_bench.log | python tools/show_bench.py)"
timeout 300 python tools/stage_profile.py cfg2 2>&1 | cut -c1-120 > gpurun_out/r2h_stage_cfg2.log; cat gpurun_out/r2h_stage_cfg2.log
