#!/bin/bash
mkdir -p gpurun_out
M2T_DEFS="AZ_NO_XSLOTS" python -m m2trans_b200.build --force > /dev/null 2>&1
SAN_TOOL=synccheck SAN_TAIL=4000 bash tools/gpu_sanitize.sh > /dev/null 2>&1
echo "--- without the extra slots:"; grep -E "ERROR SUMMARY|Barrier is located|attn_z.cu|tail_strip.cu" gpurun_out/sanitize_synccheck.log | sort | uniq -c | sort -rn | head -8
python -m m2trans_b200.build --force > /dev/null 2>&1
