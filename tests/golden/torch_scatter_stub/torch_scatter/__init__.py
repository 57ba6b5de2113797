"""Minimal pure-torch stand-in for torch-scatter 2.0.4 (absent from this image), used
ONLY by tests/golden/make_golden.py so that the unmodified reference modules import.
API subset: scatter_add / scatter_sum / scatter_mean / scatter_max / scatter_min."""
import torch


def _expand(index, src, dim):
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def _size(index, dim_size):
    if dim_size is not None:
        return int(dim_size)
    return int(index.max()) + 1 if index.numel() else 0


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    shape = list(src.shape)
    shape[dim] = _size(index, dim_size)
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return res.scatter_add_(dim, _expand(index, src, dim), src)


scatter_sum = scatter_add


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    total = scatter_add(src, index, dim, None, dim_size)
    count = scatter_add(torch.ones_like(src), index, dim, None, dim_size).clamp(min=1)
    return total / count


def _arg_reduce(src, index, dim, dim_size, largest):
    assert dim == 0
    n = _size(index, dim_size)
    shape = [n] + list(src.shape[1:])
    idx = _expand(index, src, 0)
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(0, idx, src, 'amax' if largest else 'amin', include_self=False)
    hit = src == res.gather(0, idx)
    pos = torch.arange(src.shape[0], device=src.device).view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    cand = torch.where(hit, pos, torch.full_like(pos, src.shape[0]))
    arg = torch.full(shape, src.shape[0], dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(0, idx, cand, 'amin', include_self=True)
    return res, arg


def scatter_max(src, index, dim=0, out=None, dim_size=None):
    return _arg_reduce(src, index, dim, dim_size, True)


def scatter_min(src, index, dim=0, out=None, dim_size=None):
    return _arg_reduce(src, index, dim, dim_size, False)
