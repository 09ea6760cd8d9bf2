#!/bin/bash
# compute-sanitizer memcheck over one small forward per scale (eager, both precision modes) and one TransBlock
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, types, torch
sys.path.insert(0, ".")
from m2trans_b200 import _lib
from m2trans_b200.M2Trans_network import M2Trans
from m2trans_b200.rlutrans import TransBlock
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict, synthetic_tokens, synthetic_transblock_state_dict
for scale, var in ((4, 0), (4, _lib.VAR_PRECISE_ON), (3, 0), (2, 0)):
    m = M2Trans(types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=2, kernel_variant=var)).cuda()
    m.cuda_graph = False
    m.load_state_dict(synthetic_state_dict(scale, 0, n_blocks=2))
    y = m(synthetic_input(2, 40, 72).cuda())
    torch.cuda.synchronize()
    print("ok", scale, var, tuple(y.shape), float(y.mean()))
tb = TransBlock(); tb.load_state_dict(synthetic_transblock_state_dict(0)); tb = tb.cuda()
print("ok transblock", float(tb(synthetic_tokens(2, 100).cuda()).mean()))
PY
timeout 1500 compute-sanitizer --tool ${SAN_TOOL:-memcheck} --error-exitcode 7 python /tmp/san.py 2>&1 | grep -vE "^$" | tail -${SAN_TAIL:-25} | tee gpurun_out/sanitize_${SAN_TOOL:-memcheck}.log
