"""The north-star job end to end on the GPU (generate_evp_LFB.py:439-499, trans_SV_output.py:250-301):
  * chained gate: fp32-oracle LFB features vs CUDA (bf16-operand) LFB features, BOTH through MS-TCN -> per-frame phase argmax
    agreement >= 99.9 % on a non-degenerate class histogram, top-2 margins printed;
  * sharding: videos LPT-assigned to ranks, every rank extracts its own videos into the rows of a SharedLFB -> the gathered array is
    BIT-IDENTICAL to the single-GPU result (runs on 2 GPUs as 2 processes when the box has them, else both shards on cuda:0);
  * DataParallel-wrapped call site (generate_evp_LFB.py:430-432); mit_b4_evp / mit_b5_evp; 480x854 with a ragged micro-batch."""
import os

import numpy as np
import pytest
import torch

import surgvid_b200  # noqa: F401
from oracle import evp_oracle as EO
from oracle import mstcn_oracle as MO
from surgvid_b200 import lfb
from surgvid_b200 import synthetic as S
from surgvid_b200.models import mix_transformer_evp as M
from surgvid_b200.mstcn import MultiStageModel_S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = S.EVP_CONFIGS["mit_b3_evp"]


def _evp(variant="mit_b3_evp", mode="ref_init", seed=0, dev=DEV):
    m = getattr(M, variant)()
    sd = S.synth_state_dict(S.evp_key_shapes(variant), seed=seed, mode=mode)
    m.load_state_dict(sd, strict=True)
    return m.to(dev).eval(), sd


def _oracle_feats(sd, cfg, x, seg, flow, chunk=32):
    return torch.cat([EO.evp_forward(sd, cfg, x[i:i + chunk], seg[i:i + chunk], flow[i:i + chunk]) for i in range(0, x.shape[0], chunk)])


def test_chained_phase_argmax_agreement():
    """north_star: 'per-frame phase argmax agreement >= 99.9 %' of the CHAINED path.  Frames carry a pseudo-phase pattern; MS-TCN weights
    are calibrated on 70 oracle-feature frames so that its classes respond to the phases the encoder sees (non-degenerate histogram)."""
    torch.set_num_threads(os.cpu_count() or 1)
    model, sd = _evp()
    # calibration: class-mean deviations of the oracle's features
    xc, sc, fc, pc = S.synth_phase_frames(70, seed=900, phase=torch.arange(70) % 7)
    fcal = _oracle_feats(sd, CFG, xc, sc, fc)
    center = fcal.mean(0)
    protos = torch.stack([fcal[pc == c].mean(0) for c in range(7)]) - center
    msd = S.synth_mstcn_state_dict_for_protos(protos, center)
    tcn = MultiStageModel_S(2, 8, 32, 2048, 14, True)
    tcn.load_state_dict(msd, strict=True)
    tcn = tcn.to(DEV).eval()
    lengths = [400, 333, 291]
    agree = total = 0
    hist = np.zeros(7, dtype=np.int64)
    margins, rels = [], []
    ext = lfb.LFBExtractor(model, batch_size=200, device=DEV)
    for vi, T in enumerate(lengths):
        x, seg, flow, _ = S.synth_phase_frames(T, seed=910 + vi)
        f_ref = _oracle_feats(sd, CFG, x, seg, flow)                                  # fp32 reference features
        f_gpu = ext.extract(x.pin_memory(), seg.pin_memory(), flow.pin_memory())     # host -> CUDA path -> host, as the LFB driver
        rels.append(float((f_gpu - f_ref).norm() / f_ref.norm()))
        ref = MO.mstcn_forward(msd, f_ref.unsqueeze(0).transpose(2, 1))[-1, 0]       # [14, T] reference chain, all fp32
        with torch.no_grad():
            got = tcn(f_gpu.to(DEV).unsqueeze(0).transpose(2, 1))[-1, 0].cpu()       # CUDA chain (trans_SV_output.py:271-280)
        a, b = got[:7].argmax(0), ref[:7].argmax(0)
        agree += int((a == b).sum()); total += T
        hist += np.bincount(b.numpy(), minlength=7)
        top2 = ref[:7].topk(2, dim=0).values
        margins.append(top2[0] - top2[1])
        print(f"[chain] video {vi}: T {T}, LFB rel-L2 {rels[-1]:.3e}, logits max-abs diff {float((got - ref).abs().max()):.3e} (max|ref| {float(ref.abs().max()):.2f}), "
              f"argmax agree {int((a == b).sum())}/{T}")
    margins = torch.cat(margins)
    print(f"[chain] argmax agreement {agree}/{total} = {agree / total:.5f}; class histogram {hist.tolist()}; top-2 margin median {float(margins.median()):.3f} "
          f"p1 {float(margins.kthvalue(max(1, margins.numel() // 100)).values):.4f} min {float(margins.min()):.4f}")
    assert max(rels) <= 1.5e-2
    assert (hist > 0).sum() >= 5, "degenerate class histogram makes the agreement metric vacuous"
    assert agree / total >= 0.999


# ----------------------------------------------------------------------------------------------- sharded job
_LENGTHS = [37, 12, 55, 20, 41, 9, 30, 26]


def _video(v):
    return S.synth_frames(_LENGTHS[v], seed=500 + v)


def _shard_worker(rank, world, dev, name, batch):
    torch.cuda.set_device(dev)
    model, _ = _evp(mode="stress", dev=dev)
    assign = lfb.lpt_assign(_LENGTHS, world)
    shared = lfb.SharedLFB(name, _LENGTHS, 2048)
    ext = lfb.LFBExtractor(model, batch_size=batch, device=dev)
    vids = [tuple(t.pin_memory() for t in _video(v)) for v in assign[rank]]
    ext.extract_videos(vids, outs=shared.blocks(assign[rank]))
    torch.cuda.synchronize()
    shared.close()


def test_sharded_extraction_is_bit_identical_to_single_gpu():
    model, _ = _evp(mode="stress")
    ext = lfb.LFBExtractor(model, batch_size=64, device=DEV)
    single = torch.cat(ext.extract_videos([tuple(t.pin_memory() for t in _video(v)) for v in range(len(_LENGTHS))]))
    name = f"surgvid_test_lfb_{os.getpid()}"
    shared = lfb.SharedLFB(name, _LENGTHS, 2048, create=True)
    try:
        shared.array.fill_(float("nan"))
        world = 2
        if torch.cuda.device_count() >= 2:
            import torch.multiprocessing as mp
            ctx = mp.get_context("spawn")
            procs = [ctx.Process(target=_shard_worker, args=(r, world, f"cuda:{r}", name, 48)) for r in range(world)]
            for p in procs:
                p.start()
            for p in procs:
                p.join(timeout=600)
                assert p.exitcode == 0
        else:
            for r in range(world):     # one GPU on this box: the two ranks' shards one after the other (different batch size than `single`)
                _shard_worker(r, world, DEV, name, 48)
        assert torch.equal(shared.array, single), "gathered shards differ from the single-GPU LFB"
        # consumers slice by cumulative num_each (trans_SV_output.py:56-72)
        off = np.cumsum([0] + _LENGTHS)
        assert torch.equal(shared.block(3), single[off[3]:off[4]])
    finally:
        shared.unlink()


def test_cyclic_pool_videos_equal_materialised_videos():
    model, _ = _evp(mode="stress")
    ext = lfb.LFBExtractor(model, batch_size=16, device=DEV)
    px, ps, pf = (t.pin_memory() for t in S.synth_frames(10, seed=600))
    vids = [(lfb.CyclicFrames(px, 7, 13), lfb.CyclicFrames(ps, 7, 13), lfb.CyclicFrames(pf, 7, 13)), (lfb.CyclicFrames(px, 0, 5), lfb.CyclicFrames(ps, 0, 5), lfb.CyclicFrames(pf, 0, 5))]
    dev_outs = [torch.empty((13, 2048), device=DEV), torch.empty((5, 2048), device=DEV)]
    a = ext.extract_videos(vids, device_outs=dev_outs)
    idx = [(7 + t) % 10 for t in range(13)]
    b = ext.extract_videos([(px[idx], ps[idx], pf[idx]), (px[:5], ps[:5], pf[:5])])
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(dev_outs[0].cpu(), a[0]) and torch.equal(dev_outs[1].cpu(), a[1])


def test_dataparallel_wrapped_call_site():
    """generate_evp_LFB.py:430-437: DataParallel(model).to(device); requires_grad = False; eval(); then model(x, seg, flow, return_features=True)."""
    model, sd = _evp()
    wrapped = torch.nn.DataParallel(model, device_ids=[0]).to(DEV)
    for p in wrapped.parameters():
        p.requires_grad = False
    wrapped.eval()
    x, seg, flow = S.synth_frames(4, seed=71)
    with torch.no_grad():
        out = wrapped(x.to(DEV), seg.to(DEV), flow.to(DEV), return_features=True)
        direct = model(x.to(DEV), seg.to(DEV), flow.to(DEV), return_features=True)
    assert torch.equal(out, direct)
    ref = EO.evp_forward(sd, CFG, x, seg, flow)
    assert float((out.cpu() - ref).norm() / ref.norm()) <= 1.5e-2
    # the wrapper's state_dict keys carry the 'module.' prefix the reference's checkpoints are saved with
    assert all(k.startswith("module.") for k in wrapped.state_dict())


@pytest.mark.parametrize("variant", ["mit_b4_evp", "mit_b5_evp"])
def test_deeper_variants_match_oracle(variant):
    m, sd = _evp(variant, mode="stress", seed=3)
    x, seg, flow = S.synth_frames(2, seed=78)
    ref = EO.evp_forward(sd, S.EVP_CONFIGS[variant], x, seg, flow)
    with torch.no_grad():
        out = m(x.to(DEV), seg.to(DEV), flow.to(DEV), return_features=True)
    rel = float((out.cpu() - ref).norm() / ref.norm())
    print(f"[parity] {variant}: rel-L2 {rel:.3e}")
    assert rel <= 1.5e-2 and float((out.cpu() - ref).abs().max()) <= 3e-2 * float(ref.abs().max())


def test_480x854_three_frames_ragged_microbatch_vs_oracle():
    """BASELINE configs[4] parity: N_kv = 390/390/405/405 (streamed attention), 3 frames with micro_batch 2 (plans for n = 2 and n = 1)."""
    torch.set_num_threads(os.cpu_count() or 1)
    m, sd = _evp(mode="stress")
    x, seg, flow = S.synth_frames(3, seed=321, H=480, W=854)
    ref = EO.evp_forward(sd, CFG, x, seg, flow)
    m.micro_batch = 2
    with torch.no_grad():
        out = m(x.to(DEV), seg.to(DEV), flow.to(DEV), return_features=True)
        m.micro_batch = 3
        out3 = m(x.to(DEV), seg.to(DEV), flow.to(DEV), return_features=True)
        solo = m(x[2:3].to(DEV), seg[2:3].to(DEV), flow[2:3].to(DEV), return_features=True)
    rel = float((out.cpu() - ref).norm() / ref.norm())
    print(f"[parity] 480x854 B=3 mb=2 vs oracle: rel-L2 {rel:.3e}  max-abs {float((out.cpu() - ref).abs().max()):.3e}")
    assert rel <= 1.5e-2 and float((out.cpu() - ref).abs().max()) <= 3e-2 * float(ref.abs().max())
    assert torch.equal(out, out3) and torch.equal(solo[0], out[2])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_same_process_second_gpu():
    """One process, one model object, two devices (the DataParallel replica case): the per-device native handles opt their kernels in to
    large dynamic shared memory on EACH device (the attribute is per device, not per process)."""
    model, sd = _evp(mode="stress")
    x, seg, flow = S.synth_frames(3, seed=88)
    with torch.no_grad():
        a = model(x.to("cuda:0"), seg.to("cuda:0"), flow.to("cuda:0"), return_features=True)
        model.to("cuda:1")
        with torch.cuda.device(1):
            b = model(x.to("cuda:1"), seg.to("cuda:1"), flow.to("cuda:1"), return_features=True)
            tcn = MultiStageModel_S(2, 8, 32, 2048, 14, True)
            tcn.load_state_dict(S.synth_mstcn_state_dict(mode="phase"))
            tcn = tcn.to("cuda:1").eval()
            lg1 = tcn.forward_videos(b.repeat(20, 1), [60])
        tcn = tcn.to("cuda:0")
        lg0 = tcn.forward_videos(a.repeat(20, 1), [60])
    model.to("cuda:0")
    assert torch.equal(a.cpu(), b.cpu()) and torch.equal(lg0.cpu(), lg1.cpu())
