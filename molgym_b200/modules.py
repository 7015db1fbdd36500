"""The `molgym.modules` surface (molgym/modules.py:8-50) kept importable under the same names: MLP, masked_softmax,
to_one_hot, init_layer.  Plain torch; the hot path does not run through these (the CUDA kernels carry their own MLPs),
they serve the rollout-mode sampling code and user code that imports them."""
from typing import Tuple

import torch
import torch.nn as nn


def to_one_hot(indices: torch.Tensor, num_classes: int, device=None) -> torch.Tensor:
    """modules.py:8-23 — indices [..., 1] -> one-hot [..., num_classes]; out-of-range raises RuntimeError."""
    shape = indices.shape[:-1] + (num_classes, )
    oh = torch.zeros(shape, device=device if device is not None else indices.device).view(shape)
    if indices.numel() and (int(indices.min()) < 0 or int(indices.max()) >= num_classes):
        raise RuntimeError(f'index out of range for one-hot with {num_classes} classes')
    return oh.scatter_(dim=-1, index=indices, value=1)


def masked_softmax(logits: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """modules.py:26-27 — torch_scatter.composite.scatter_softmax(logits, mask.long()) * mask, i.e. a softmax inside the
    mask==1 group (and, multiplied away, one inside the mask==0 group)."""
    mask = mask.to(torch.bool)
    neg = torch.finfo(logits.dtype).min
    mx = torch.where(mask, logits, torch.full_like(logits, neg)).max(dim=-1, keepdim=True)[0]
    e = torch.where(mask, (logits - mx).exp(), torch.zeros_like(logits))
    return e / (e.sum(dim=-1, keepdim=True) + 1e-12) * mask


def init_layer(layer: nn.Linear, w_scale=1.0) -> nn.Linear:
    nn.init.orthogonal_(layer.weight.data)
    layer.weight.data.mul_(w_scale)
    nn.init.constant_(layer.bias.data, 0)
    return layer


class MLP(nn.Module):
    """modules.py:37-50."""

    def __init__(self, input_dim: int, output_dims: Tuple[int, ...], gate=torch.relu):
        super().__init__()
        dims = (input_dim, ) + tuple(output_dims)
        self.layers = nn.ModuleList([init_layer(nn.Linear(a, b)) for a, b in zip(dims[:-1], dims[1:])])
        self.gate = gate

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for layer in self.layers[:-1]:
            x = self.gate(layer(x))
        return self.layers[-1](x)
