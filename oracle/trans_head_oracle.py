"""CPU restatement of the reference-defined part of `Transformer.original_forward` (adapter_transformer.py:329-349).

TEST INFRASTRUCTURE ONLY.  The inner `Transformer2_3_1` is absent from /root/reference; the two tensors handed to it
(the causal windows `inputs` and the decoder query `feas`) are restated from the reference's own lines, the inner module itself from
the published upstream architecture (second half of this file, PARITY UNPINNED).  The reference module itself cannot be
imported (its file imports the missing module at line 9), hence PARITY UNPINNED against a live reference; the restatement
follows the source line by line (the only change: `.cuda()` at :338 dropped).
"""
from __future__ import annotations

import torch


def original_forward_inputs(x: torch.Tensor, long_feature: torch.Tensor, fc_weight: torch.Tensor, len_q: int):
    """x [1, C, T] (last MS-TCN stage), long_feature [1, T, f_dim], fc_weight [C, f_dim] -> (inputs [T, len_q, C], feas [T, 1, C])."""
    num_classes = fc_weight.shape[0]
    out_features = x.transpose(1, 2)                                     # :330
    inputs = []
    for i in range(out_features.size(1)):                                # :336-343
        if i < len_q - 1:
            pad = torch.zeros((1, len_q - 1 - i, num_classes))
            inp = torch.cat([pad, out_features[:, 0:i + 1]], dim=1)
        else:
            inp = out_features[:, i - len_q + 1:i + 1]
        inputs.append(inp)
    inputs = torch.stack(inputs, dim=0).squeeze(1)                       # :344
    feas = torch.tanh(torch.nn.functional.linear(long_feature, fc_weight).transpose(0, 1))  # :347-348
    return inputs, feas


# ----------------------------------------------------------------------------------------------------------------------------------
# Inner module `Transformer2_3_1` — PARITY UNPINNED, and doubly so: the file `transformer2_3_1.py` is absent from /root/reference
# (adapter_transformer.py:9 imports it), no requirements file pins its origin, and nothing in this image contains it.
# What the reference fixes is the constructor call (adapter_transformer.py:317-325: d_model = out_features = 14, d_ff = mstcn_f_maps,
# d_k = d_v = min(64, mstcn_f_maps), n_layers = 1, n_heads = 4, len_q = sequence_length) and the call
# `transformer(inputs [T, len_q, d_model], feas [T, 1, d_model]) -> [T, 1, d_model]` (:348).  The restatement below follows the PUBLISHED
# architecture of the upstream project the class name comes from (Trans-SVNet, Gao et al., MICCAI 2021, its `transformer2_3_1.py`:
# a one-layer encoder over the len_q-frame window of temporal embeddings — multi-head self-attention and a position-wise feed-forward
# net, each followed by residual + LayerNorm (post-LN), no positional encoding, no mask — and a one-layer decoder whose single query,
# the spatial embedding of the current frame, cross-attends the encoder output, followed by the same feed-forward block).  Parameter
# names (`W_Q`, `W_K`, `W_V`, `fc`, `layer_norm`, `pos_ffn.fc1/fc2`) are this repo's choice; biases are used when present in the
# state_dict and treated as zero otherwise, so both the biased and the bias-free variant of the upstream tutorial code can be loaded.
def _lin(sd, prefix, x):
    w = sd[prefix + ".weight"]
    b = sd.get(prefix + ".bias")
    return torch.nn.functional.linear(x, w, b)


def _mha(sd, prefix, q_in, k_in, v_in, n_heads, d_k, d_v):
    """residual + LayerNorm(fc(softmax(Q K^T / sqrt(d_k)) V)); q_in [B, Lq, D], k_in = v_in [B, Lk, D]."""
    B, Lq, D = q_in.shape
    Lk = k_in.shape[1]
    q = _lin(sd, prefix + ".W_Q", q_in).view(B, Lq, n_heads, d_k).transpose(1, 2)
    k = _lin(sd, prefix + ".W_K", k_in).view(B, Lk, n_heads, d_k).transpose(1, 2)
    v = _lin(sd, prefix + ".W_V", v_in).view(B, Lk, n_heads, d_v).transpose(1, 2)
    attn = torch.softmax(q @ k.transpose(-1, -2) / (d_k ** 0.5), dim=-1)
    ctx = (attn @ v).transpose(1, 2).reshape(B, Lq, n_heads * d_v)
    out = _lin(sd, prefix + ".fc", ctx) + q_in
    return torch.nn.functional.layer_norm(out, (D,), sd[prefix + ".layer_norm.weight"], sd[prefix + ".layer_norm.bias"], 1e-5)


def _ffn(sd, prefix, x):
    D = x.shape[-1]
    out = _lin(sd, prefix + ".fc2", torch.relu(_lin(sd, prefix + ".fc1", x))) + x
    return torch.nn.functional.layer_norm(out, (D,), sd[prefix + ".layer_norm.weight"], sd[prefix + ".layer_norm.bias"], 1e-5)


def transformer2_3_1_forward(sd, enc_inputs: torch.Tensor, dec_inputs: torch.Tensor, n_heads: int, d_k: int, d_v: int, n_layers: int = 1):
    """sd: state_dict of surgvid_b200.trans_head.Transformer2_3_1; enc_inputs [T, len_q, d_model], dec_inputs [T, 1, d_model]
    -> [T, 1, d_model]."""
    enc = enc_inputs
    for l in range(n_layers):
        p = f"encoder.layers.{l}"
        enc = _mha(sd, p + ".enc_self_attn", enc, enc, enc, n_heads, d_k, d_v)
        enc = _ffn(sd, p + ".pos_ffn", enc)
    dec = dec_inputs
    for l in range(n_layers):
        p = f"decoder.layers.{l}"
        dec = _mha(sd, p + ".dec_enc_attn", dec, enc, enc, n_heads, d_k, d_v)
        dec = _ffn(sd, p + ".pos_ffn", dec)
    return dec
