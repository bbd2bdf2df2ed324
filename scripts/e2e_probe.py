"""Where does the end-to-end gap come from?  Same loop as LFBExtractor.extract_videos with pieces switched off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, surgvid_b200
from surgvid_b200 import synthetic as S, lfb
from surgvid_b200.models.mix_transformer_evp import mit_b3_evp
dev = "cuda:0"
m = mit_b3_evp(); m.load_state_dict(S.synth_state_dict(S.evp_key_shapes("mit_b3_evp"), seed=0, mode="ref_init")); m = m.to(dev).eval(); m.micro_batch = 800
T = 2300
x, seg, flow = S.synth_frames(T, seed=1)
xh, sh, fh = x.pin_memory(), seg.pin_memory(), flow.pin_memory()
xd, sd, fd = x.to(dev), seg.to(dev), flow.to(dev)
outs = [torch.empty((T, 2048)).pin_memory() for _ in range(3)]

def timed(fn, n=2):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n

def device_only():
    with torch.no_grad():
        for _ in range(3):
            for b0 in range(0, T, 800):
                m(xd[b0:b0 + 800], sd[b0:b0 + 800], fd[b0:b0 + 800], return_features=True)
dt = timed(device_only); print(f"device-resident, 3 videos, batches of 800: {3*T/dt:8.0f} frames/s")
ex = lfb.LFBExtractor(m, batch_size=800)
dt = timed(lambda: ex.extract_videos([(xh, sh, fh)] * 3, outs=outs)); print(f"extract_videos (H2D overlapped):          {3*T/dt:8.0f} frames/s")
ex2 = lfb.LFBExtractor(m, batch_size=800, ramp_start=800)
dt = timed(lambda: ex2.extract_videos([(xh, sh, fh)] * 3, outs=outs)); print(f"extract_videos, no ramp:                   {3*T/dt:8.0f} frames/s")
# copies only
cs = torch.cuda.Stream()
def copies_only():
    with torch.cuda.stream(cs):
        for _ in range(3):
            xd.copy_(xh, non_blocking=True); sd.copy_(sh, non_blocking=True); fd.copy_(fh, non_blocking=True)
    cs.synchronize()
dt = timed(copies_only); print(f"H2D copies alone: {3*3.7118/dt:6.1f} GB/s -> {3*T/dt:8.0f} frames/s equivalent")
# compute while an unrelated H2D stream runs flat out
def both():
    with torch.cuda.stream(cs):
        for _ in range(3):
            xd2.copy_(xh, non_blocking=True); sd2.copy_(sh, non_blocking=True); fd2.copy_(fh, non_blocking=True)
    device_only()
    cs.synchronize()
xd2, sd2, fd2 = torch.empty_like(xd), torch.empty_like(sd), torch.empty_like(fd)
dt = timed(both); print(f"device-resident compute with a concurrent H2D stream: {3*T/dt:8.0f} frames/s")
