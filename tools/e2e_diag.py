import sys, time, types, torch
sys.path.insert(0, "/root/repo")
from m2trans_b200.M2Trans_network import M2Trans
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict
m = M2Trans(types.SimpleNamespace(scale=4, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8)).cuda()
m.load_state_dict(synthetic_state_dict(4, 0))
B,H,W=16,128,128
xh=[synthetic_input(B,H,W,seed=i).pin_memory() for i in range(3)]
yh=[torch.empty(B,3,512,512).pin_memory() for _ in range(2)]
# pure copy speeds
d=torch.empty(B,3,512,512,device='cuda')
for _ in range(3): yh[0].copy_(d, non_blocking=True)
torch.cuda.synchronize()
ts=[]
for i in range(20):
    t0=time.perf_counter(); yh[i%2].copy_(d, non_blocking=True); torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
print("D2H 50MB ms:", [round(t,2) for t in ts])
xd=[torch.empty(B,3,H,W,device='cuda') for _ in range(2)]
s_in,s_out,s_main=torch.cuda.Stream(),torch.cuda.Stream(),torch.cuda.current_stream()
ev_in=[torch.cuda.Event() for _ in range(2)]; ev_done=[torch.cuda.Event() for _ in range(2)]
def step(i):
    k=i%2
    with torch.cuda.stream(s_in):
        s_in.wait_event(ev_done[k]); xd[k].copy_(xh[i%3], non_blocking=True); ev_in[k].record(s_in)
    s_main.wait_event(ev_in[k]); y=m(xd[k]); ev_done[k].record(s_main)
    with torch.cuda.stream(s_out):
        s_out.wait_event(ev_done[k]); yh[k].copy_(y, non_blocking=True); y.record_stream(s_out)
for i in range(5): step(i)
torch.cuda.synchronize()
marks=[]
t0=time.perf_counter()
for i in range(40):
    step(i); marks.append(time.perf_counter()-t0)
torch.cuda.synchronize(); tot=time.perf_counter()-t0
print("e2e ms/step", tot/40*1e3, "host enqueue ms/step", marks[-1]/40*1e3)
