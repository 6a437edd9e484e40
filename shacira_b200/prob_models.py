"""Factorized-density probability model, mirror of `wisp/models/prob_models/bit_estimator.py`
(Bitparm :9-44, BitEstimator :46-65): same parameter names (`f1..f4` x `h,b,a`, shape [1, C]).

`forward()` is the PyTorch definition (used by `size(use_prob_model=True)` on a handful of
unique symbols and by the parity tests). The training-time bit-rate loss does not call it:
`LatentGrid.ent_loss` packs the parameters with `packed_params()` and runs the fused CUDA
kernel (value + gradients in one pass). `BitEstimatorN` (bit_estimatorN.py) is never
constructed by the grids (latent_grid.py:119) and is out of scope.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Bitparm(nn.Module):
    def __init__(self, channel, is_symmetric=False, is_unimodal=False, final=False):
        super().__init__()
        self.final = final
        self.is_unimodal = is_unimodal
        self.h = nn.Parameter(torch.empty(1, channel).normal_(0, 0.01))
        if is_symmetric:
            self.b = nn.Parameter(torch.zeros(1, channel), requires_grad=False)
        else:
            self.b = nn.Parameter(torch.empty(1, channel).normal_(0, 0.01))
        self.a = None if final else nn.Parameter(torch.empty(1, channel).normal_(0, 0.01))

    def forward(self, x, single_channel=None):
        pick = (lambda p: p[:, single_channel]) if single_channel is not None else (lambda p: p)
        h, b = pick(self.h), pick(self.b)
        x = x * F.softplus(h) + b
        if self.final:
            return torch.sigmoid(x)
        a = pick(self.a)
        if self.is_unimodal:
            a = torch.abs(a)
        return x + torch.tanh(x) * torch.tanh(a)


class BitEstimator(nn.Module):
    def __init__(self, channel, is_symmetric=False, is_unimodal=False, num_layers=4):
        super().__init__()
        self.num_layers = num_layers
        self.channel = channel
        self.is_unimodal = is_unimodal
        self.f1 = Bitparm(channel, is_symmetric, is_unimodal)
        self.f2 = Bitparm(channel, is_symmetric, is_unimodal)
        self.f3 = Bitparm(channel, is_symmetric, is_unimodal)
        self.f4 = Bitparm(channel, is_symmetric, is_unimodal, final=True)

    def forward(self, x, single_channel=None):
        if self.num_layers > 1:
            x = self.f1(x, single_channel)
        if self.num_layers > 2:
            x = self.f2(x, single_channel)
        if self.num_layers > 3:
            x = self.f3(x, single_channel)
        return self.f4(x, single_channel)

    def packed_params(self):
        """[4, 3, C] = {f1..f4} x {h, b, a} for the fused kernel (differentiable stack)."""
        rows = []
        for f in (self.f1, self.f2, self.f3, self.f4):
            a = f.a if f.a is not None else torch.zeros_like(f.h)
            if self.is_unimodal and f.a is not None:
                a = torch.abs(a)
            rows.append(torch.cat((f.h, f.b, a), dim=0))
        return torch.stack(rows, dim=0)
