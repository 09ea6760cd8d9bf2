#!/bin/bash
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from m2trans_b200 import _lib
lib=_lib.load()
B,hp,wp=16,128,128
Y=torch.randn(B,hp,wp,64,device='cuda').half(); Xin=torch.randn(B,hp,wp,64,device='cuda'); Xout=torch.empty_like(Xin)
ffw=(torch.randn(9,64,64,device='cuda')*0.05).half(); ffb=torch.randn(64,device='cuda'); stats=torch.zeros(B,64,2,dtype=torch.float64,device='cuda')
flush=torch.empty(192*1024*1024,dtype=torch.float32,device='cuda')
ts=[]
for i in range(15):
    flush.zero_(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); _lib.check(lib.m2t_stage_ffconv(0,Y.data_ptr(),ffw.data_ptr(),ffb.data_ptr(),Xin.data_ptr(),Xout.data_ptr(),stats.data_ptr(),B,hp,wp,None),'c'); b.record(); torch.cuda.synchronize()
    if i>=3: ts.append(a.elapsed_time(b)*1e3)
print('ffconv us: min %.1f median %.1f'%(min(ts), sorted(ts)[len(ts)//2]))
PY
