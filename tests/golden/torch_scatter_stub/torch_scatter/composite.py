"""torch_scatter.composite.scatter_softmax (2.0.4): exp(src - segmax) / (segsum + eps)."""
import torch

from . import _expand, scatter_add


def scatter_softmax(src, index, dim=0, eps=1e-12):
    n = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = n
    idx = _expand(index, src, dim)
    seg_max = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
    seg_max = seg_max.scatter_reduce(dim, idx, src, 'amax', include_self=True)
    ex = (src - seg_max.gather(dim, idx)).exp()
    seg_sum = scatter_add(ex, index, dim, None, n)
    return ex / (seg_sum.gather(dim, idx) + eps)
