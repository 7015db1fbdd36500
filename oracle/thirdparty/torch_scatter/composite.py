"""ORACLE / TEST INFRASTRUCTURE — restated `torch_scatter.composite.scatter_softmax` (2.0.5):
softmax within each index group along `dim` (group max subtracted, eps added to the group sum).
Behaviour pinned by the reference's tests/test_modules.py:31-47."""
import torch


def scatter_softmax(src, index, dim=-1, eps=1e-12):
    if not torch.is_floating_point(src):
        raise ValueError('`scatter_softmax` can only be computed over tensors with floating point data types.')
    num_groups = int(index.max().item()) + 1 if index.numel() > 0 else 0
    shape = list(src.shape)
    shape[dim] = num_groups
    neg_inf = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
    max_per_group = neg_inf.scatter_reduce(dim, index, src.detach(), reduce='amax', include_self=True)
    recentered = src - max_per_group.gather(dim, index)
    exp = recentered.exp()
    sum_per_group = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(dim, index, exp)
    norm = (sum_per_group + eps).gather(dim, index)
    return exp / norm
