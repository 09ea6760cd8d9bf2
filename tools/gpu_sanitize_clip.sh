#!/bin/bash
# compute-sanitizer over one MedCLIP image pass (batch 2) and the stage entry points
mkdir -p gpurun_out
cat > /tmp/san_clip.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from m2trans_b200.medclip_image import MedCLIPVisionModelViT, synthetic_state_dict
t = MedCLIPVisionModelViT(); t.load_state_dict(synthetic_state_dict(0), strict=False); t = t.cuda()
e, l = t.encode_image(torch.rand(2, 3, 100, 140, device="cuda"), torch.randn(512, device="cuda"))
torch.cuda.synchronize()
print("ok clip", tuple(e.shape), float(l.mean()))
PY
timeout 1200 compute-sanitizer --tool ${SAN_TOOL:-memcheck} --error-exitcode 7 python /tmp/san_clip.py 2>&1 | grep -vE "^$" | tail -${SAN_TAIL:-15} | tee gpurun_out/sanitize_clip_${SAN_TOOL:-memcheck}.log
