#!/bin/bash
# evidence run (cfg2): launch list of two forwards, then ncu --set full of one launch of every kernel class of the forward
# (the kernel filter drops the one-time weight-packing kernels; skip counts land in the first timed forward)
mkdir -p gpurun_out
K='regex:attn|ffconv|tail|branch_prep|head_conv'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 153 -c 102 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_launches.log 2>&1
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o gpurun_out/r02_$1 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_ncu_$1.log 2>&1
}
cap attn_z attn_z_kernel 72 3
cap attn16 attn16_qkv 24 1
cap ffconv ffconv_umma 24 1
cap prep branch_prep_all 24 1
cap head head_conv 3 1
cap tail_strip tail_strip 3 1
cap tail_up tail_up_umma 3 1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
