"""The part of the reference's Trans-SVNet head that the reference itself defines.

`adapter_transformer.Transformer` (adapter_transformer.py:290-352) wraps an inner `Transformer2_3_1` module whose source is
NOT in the reference (it is imported from a file that does not exist there), so that module cannot be restated or pinned.
What `Transformer.original_forward` does around it is fully specified and is reproduced here on the GPU:

  * `inputs[t] = out_features[t-len_q+1 .. t]` with zero left padding — a `[T, len_q, 14]` tensor of causal windows of the
    MS-TCN logits (adapter_transformer.py:335-344; the reference builds it with a Python loop and `.cuda()` calls);
  * `feas = tanh(fc(long_feature))` with `fc = Linear(f_dim, out_features, bias=False)` (adapter_transformer.py:325,348),
    computed by the MS-TCN stage-1 projection kernel in the SAME pass over the `[T, 2048]` features (14 extra columns).

`TransformerInputs` has the constructor arguments and the `fc.weight` state_dict key of the reference class; the inner module
is whatever the user supplies (`original_forward(..., transformer=module)`), called exactly as the reference calls it.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .mstcn import MultiStageModel_S


class TransformerInputs(nn.Module):
    def __init__(self, mstcn_f_maps, mstcn_f_dim, out_features, len_q, **kwargs):
        super().__init__()
        self.num_f_maps, self.dim, self.num_classes, self.len_q = mstcn_f_maps, mstcn_f_dim, out_features, len_q
        self.fc = nn.Linear(mstcn_f_dim, out_features, bias=False)  # adapter_transformer.py:325
        self._attached_to = None

    def attach(self, mstcn_model: MultiStageModel_S):
        """Pack `fc.weight` next to the MS-TCN stage-1 projection of `mstcn_model` (call again after changing fc.weight)."""
        mstcn_model.set_query_head(self.fc.weight)
        self._attached_to = mstcn_model
        return self

    @torch.no_grad()
    def prepare(self, mstcn_model: MultiStageModel_S, long_feature: torch.Tensor, lengths: Sequence[int]):
        """long_feature [sum(lengths), f_dim] fp32 CUDA, videos concatenated ->
        (logits [stages, C, T], inputs [T, len_q, C], feas [T, 1, C]) — the arguments of `self.transformer(inputs, feas)`."""
        if self._attached_to is not mstcn_model:
            self.attach(mstcn_model)
        logits, query = mstcn_model.forward_videos_query(long_feature, lengths)
        inputs = ops.causal_windows(logits[-1], lengths, self.len_q)
        return logits, inputs, query.unsqueeze(1)

    @torch.no_grad()
    def original_forward(self, x: torch.Tensor, long_feature: torch.Tensor, transformer: Optional[nn.Module] = None):
        """Reference signature (adapter_transformer.py:329): x = last-stage MS-TCN output [1, C, T], long_feature [1, T, f_dim].
        Returns `transformer(inputs, feas)` when the inner module is supplied, else the pair (inputs, feas)."""
        if not x.is_cuda:
            raise RuntimeError("surgvid_b200 has no CPU path: move the inputs to a CUDA (sm_100a) device")
        T = x.shape[2]
        inputs = ops.causal_windows(x[0].to(torch.float32).contiguous(), [T], self.len_q)
        lf = long_feature.reshape(T, self.dim).to(torch.float32).contiguous()
        if self._attached_to is None:
            raise RuntimeError("call attach(mstcn_model) first: the fc projection runs inside the MS-TCN stage-1 kernel")
        _, q = self._attached_to.forward_videos_query(lf, [T])
        feas = q.unsqueeze(1)
        return transformer(inputs, feas) if transformer is not None else (inputs, feas)
