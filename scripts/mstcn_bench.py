"""MS-TCN timing over the 80 Cholec80-shaped synthetic feature sequences (BASELINE.json configs[3]) in one batched call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import surgvid_b200  # noqa
from surgvid_b200 import synthetic as S
from surgvid_b200.mstcn import MultiStageModel_S
dev = "cuda:0"
L = [int(v) for v in S.cholec80_video_lengths()]
T = sum(L)
m = MultiStageModel_S(2, 8, 32, 2048, 14, True)
m.load_state_dict(S.synth_mstcn_state_dict(mode="phase")); m = m.to(dev).eval()
feats = torch.rand(T, 2048, device=dev) * 0.57
for _ in range(3): m.forward_videos(feats, L)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = int(os.environ.get("REPS", "10"))
e0.record()
for _ in range(reps): out = m.forward_videos(feats, L)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"MS-TCN 80 videos, {T} frames: {ms:.3f} ms  -> {T/ms*1e3/1e6:.2f} M frames/s, feature read {T*8192/ms/1e6:.1f} GB/s ({T*8192/ms/1e6/6446.9:.3f} of HBM peak)")
if os.environ.get("HEAD", "1") == "1":
    from surgvid_b200.trans_head import Transformer
    head = Transformer(32, 2048, 14, 30).to(dev).eval()
    for _ in range(2): head.forward_videos(m, feats, L)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): lg, o = head.forward_videos(m, feats, L)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / reps
    print(f"MS-TCN + Trans-SVNet head (query fused into the projection, windows never materialised), {T} frames: {ms2:.3f} ms (head alone {ms2 - ms:.3f} ms)")
