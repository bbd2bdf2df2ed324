"""CPU restatement of the reference-defined part of `Transformer.original_forward` (adapter_transformer.py:329-349).

TEST INFRASTRUCTURE ONLY.  The inner `Transformer2_3_1` is absent from /root/reference, so only the two tensors
handed to it are restated: the causal windows `inputs` and the decoder query `feas`.  The reference module itself cannot be
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
