import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, surgvid_b200
from surgvid_b200 import synthetic as S, lfb
from surgvid_b200.models.mix_transformer_evp import mit_b3_evp
dev="cuda:0"
m=mit_b3_evp(); m.load_state_dict(S.synth_state_dict(S.evp_key_shapes("mit_b3_evp"),seed=0,mode="ref_init")); m=m.to(dev).eval(); m.micro_batch=800
T=2300
x,seg,flow=S.synth_frames(T,seed=1)
xh,sh,fh=x.pin_memory(),seg.pin_memory(),flow.pin_memory()
out=torch.empty((T,2048)).pin_memory()
for rs in (200, 100, 50, 25, 400, 800):
    ex=lfb.LFBExtractor(m,batch_size=800,ramp_start=rs)
    for _ in range(2): ex.extract(xh,sh,fh,out=out)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(3): ex.extract(xh,sh,fh,out=out)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
    print(f"ramp_start {rs:4d}: {T/dt:8.0f} frames/s  ({dt*1e3:.1f} ms)  schedule {[n for _,n in ex._schedule(T)]}", flush=True)
