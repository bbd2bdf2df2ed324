"""Reference-defined inputs of the Trans-SVNet head: causal windows + fused tanh(fc(LFB)) query."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _models(f_maps=32):
    from surgvid_b200 import synthetic
    from surgvid_b200.mstcn import MultiStageModel_S
    from surgvid_b200.trans_head import TransformerInputs
    tcn = MultiStageModel_S(2, 8, f_maps, 2048, 14, True)
    sd = synthetic.synth_mstcn_state_dict(2, 8, f_maps, 2048, 14, seed=3, mode="phase")
    tcn.load_state_dict(sd)
    tcn = tcn.cuda().eval()
    head = TransformerInputs(f_maps, 2048, 14, 30)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        head.fc.weight.copy_(torch.randn(14, 2048, generator=g) * 0.05)
    return tcn, head


@pytest.mark.parametrize("f_maps", [32, 64])
def test_query_and_windows_match_oracle(f_maps):
    from oracle import trans_head_oracle as tho
    from oracle import mstcn_oracle
    from surgvid_b200 import synthetic
    tcn, head = _models(f_maps)
    lengths = [45, 700, 29, 333]
    T = sum(lengths)
    lfb = synthetic.synth_lfb_features(T, seed=9).cuda()
    plain = tcn.forward_videos(lfb, lengths)                 # before the head is attached
    logits, inputs, feas = head.prepare(tcn, lfb, lengths)
    assert inputs.shape == (T, 30, 14) and feas.shape == (T, 1, 14)
    assert torch.equal(logits, plain)                        # the 14 extra columns do not disturb the MS-TCN result
    sd = {k: v.detach().cpu() for k, v in tcn.state_dict().items()}
    o = 0
    for Tv in lengths:
        x = lfb[o:o + Tv].cpu()
        ref_logits = mstcn_oracle.mstcn_forward(sd, x.t().unsqueeze(0))          # [stages, 1, 14, Tv]
        want_in, want_q = tho.original_forward_inputs(ref_logits[-1], x.unsqueeze(0), head.fc.weight.detach().cpu(), 30)
        got_in, got_q = inputs[o:o + Tv].cpu(), feas[o:o + Tv].cpu()
        # windows are a gather of the device logits: exact against the device logits, 1e-4-close to the fp32 oracle's
        dev_in, _ = tho.original_forward_inputs(logits[-1][:, o:o + Tv].cpu().unsqueeze(0), x.unsqueeze(0), head.fc.weight.detach().cpu(), 30)
        assert torch.equal(got_in, dev_in)
        assert float((got_in - want_in).abs().max()) <= 1e-4 * max(1.0, float(want_in.abs().max()))
        assert float((got_q - want_q).abs().max()) <= 2e-5   # 3xTF32 projection + tanh vs fp32
        o += Tv


def test_original_forward_signature_and_detach():
    from surgvid_b200 import synthetic
    tcn, head = _models(32)
    T = 211
    lfb = synthetic.synth_lfb_features(T, seed=1).cuda()
    head.attach(tcn)
    out = tcn(lfb.t().unsqueeze(0))                          # [2, 1, 14, T]  (trans_SV_output.py:279)
    inputs, feas = head.original_forward(out[-1], lfb.unsqueeze(0))
    assert inputs.shape == (T, 30, 14) and feas.shape == (T, 1, 14)
    assert torch.equal(inputs[:, -1, :], out[-1, 0].t())     # the newest frame of each window is the frame itself
    assert float(inputs[0, :29].abs().max()) == 0.0
    # a user-supplied inner module is called as the reference calls it
    class Inner(torch.nn.Module):
        def forward(self, enc_inputs, dec_inputs):
            return dec_inputs + enc_inputs[:, -1:, :]
    y = head.original_forward(out[-1], lfb.unsqueeze(0), transformer=Inner())
    assert y.shape == (T, 1, 14)
    tcn.set_query_head(None)
    with pytest.raises(RuntimeError):
        tcn.forward_videos_query(lfb, [T])


# ---- the inner module Transformer2_3_1 (absent from the reference tree: PARITY UNPINNED, checked against the restatement of the
# published upstream architecture in oracle/trans_head_oracle.py)
def _inner(bias=True, seed=11):
    from surgvid_b200.trans_head import Transformer2_3_1
    m = Transformer2_3_1(d_model=14, d_ff=32, d_k=32, d_v=32, n_layers=1, n_heads=4, len_q=30, bias=bias)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("layer_norm.weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif n.endswith("bias"):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * (1.5 / p.shape[1] ** 0.5))
    return m.cuda().eval()


@pytest.mark.parametrize("bias", [True, False])
def test_inner_transformer_matches_oracle(bias):
    from oracle import trans_head_oracle as tho
    m = _inner(bias)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    lengths = [1, 29, 30, 31, 200, 7]
    T = sum(lengths)
    logits = 3.0 * torch.randn(14, T, generator=g)
    query = torch.tanh(torch.randn(T, 14, generator=g))
    out = m.forward_fused(logits.cuda(), query.cuda(), lengths)          # [T, 1, 14]
    torch.cuda.synchronize()
    assert out.shape == (T, 1, 14)
    o = 0
    worst = 0.0
    for Tv in lengths:
        x = logits[:, o:o + Tv].unsqueeze(0)
        inputs, _ = tho.original_forward_inputs(x, torch.zeros(1, Tv, 8), torch.zeros(14, 8), 30)   # the reference's window loop
        ref = tho.transformer2_3_1_forward(sd, inputs, query[o:o + Tv].unsqueeze(1), 4, 32, 32)
        err = float((out[o:o + Tv].cpu() - ref).abs().max())
        worst = max(worst, err)
        # the reference's own call form: windows + query of ONE video
        got2 = m(inputs.cuda(), query[o:o + Tv].unsqueeze(1).cuda())
        assert torch.equal(got2, out[o:o + Tv])
        o += Tv
    print(f"[parity] Transformer2_3_1 (bias={bias}) vs unpinned oracle: max-abs {worst:.3e}")
    assert worst <= 2e-4


def test_full_head_forward_videos_and_reference_call_site():
    """trans_SV_output.py:268-301: per video, out = mstcn(video_fe)[-1]; output = transformer(out.detach(), long_feature)."""
    from oracle import trans_head_oracle as tho
    from surgvid_b200 import synthetic
    from surgvid_b200.mstcn import MultiStageModel_S
    from surgvid_b200.trans_head import Transformer
    tcn = MultiStageModel_S(2, 8, 32, 2048, 14, True)
    tcn.load_state_dict(synthetic.synth_mstcn_state_dict(2, 8, 32, 2048, 14, seed=3, mode="phase"))
    tcn = tcn.cuda().eval()
    head = Transformer(32, 2048, 14, 30)
    inner = _inner(True, seed=21)
    head.transformer.load_state_dict(inner.state_dict())
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        head.fc.weight.copy_(torch.randn(14, 2048, generator=g) * 0.05)
    head = head.cuda().eval()
    lengths = [45, 300, 29]
    T = sum(lengths)
    lfb = synthetic.synth_lfb_features(T, seed=9).cuda()
    logits, out = head.forward_videos(tcn, lfb, lengths)
    assert logits.shape == (2, 14, T) and out.shape == (T, 1, 14)
    sd = {k: v.detach().cpu() for k, v in head.transformer.state_dict().items()}
    o = 0
    for Tv in lengths:
        # reference call site, one video at a time
        video_fe = lfb[o:o + Tv].unsqueeze(0).transpose(2, 1)
        x = tcn(video_fe)[-1]                                            # [1, 14, Tv]
        y = head(x.detach(), lfb[o:o + Tv].unsqueeze(0))                  # [Tv, 1, 14]
        assert torch.equal(y, out[o:o + Tv])
        inputs, feas = tho.original_forward_inputs(x.cpu(), lfb[o:o + Tv].cpu().unsqueeze(0), head.fc.weight.detach().cpu(), 30)
        ref = tho.transformer2_3_1_forward(sd, inputs, feas, 4, 32, 32)
        assert float((y.cpu() - ref).abs().max()) <= 3e-4
        o += Tv
