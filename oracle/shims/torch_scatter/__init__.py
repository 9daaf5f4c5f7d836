"""Stand-in for torch-scatter (oracle infrastructure; see oracle/shims/README.md).

Restates torch_scatter.scatter for the reduce modes the reference calls
(/root/reference/cad_recognition/architecture3cc_rpn_gp_iter2.py:67,122 and PyG aggregate).
"""
import torch


def _bcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    while index.dim() < src.dim():
        index = index.unsqueeze(-1)
    return index.expand_as(src), dim


class _ScatterMax(torch.autograd.Function):
    """max with a single arg index per output element (first occurrence), as torch_scatter's CPU kernel."""

    @staticmethod
    def forward(ctx, src, index, dim_size):
        # src [M, F...] reduced over dim 0 with 1-D index
        M = src.shape[0]
        flat = src.reshape(M, -1)
        F_ = flat.shape[1]
        out = flat.new_full((dim_size, F_), float('-inf'))
        idx2 = index.view(-1, 1).expand(M, F_)
        out = out.scatter_reduce(0, idx2, flat, reduce='amax', include_self=True)
        # first row attaining the max
        is_max = flat == out.gather(0, idx2)
        rows = torch.arange(M, device=src.device).view(-1, 1).expand(M, F_)
        cand = torch.where(is_max, rows, torch.full_like(rows, M))
        arg = torch.full((dim_size, F_), M, dtype=torch.long, device=src.device)
        arg = arg.scatter_reduce(0, idx2, cand, reduce='amin', include_self=True)
        empty = arg == M
        out = torch.where(empty, torch.zeros_like(out), out)
        ctx.save_for_backward(arg, empty)
        ctx.shape = src.shape
        return out.reshape((dim_size,) + src.shape[1:])

    @staticmethod
    def backward(ctx, g):
        arg, empty = ctx.saved_tensors
        M = ctx.shape[0]
        g2 = g.reshape(arg.shape)
        g2 = torch.where(empty, torch.zeros_like(g2), g2)
        out = g2.new_zeros((M + 1, arg.shape[1]))
        out.scatter_(0, arg, g2)
        return out[:M].reshape(ctx.shape), None, None


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    assert out is None
    if dim < 0:
        dim = src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    if reduce == 'max':
        assert dim == 0 and index.dim() == 1
        return _ScatterMax.apply(src, index, dim_size)
    idx, dim = _bcast(index, src, dim)
    shape = list(src.shape)
    shape[dim] = dim_size
    res = src.new_zeros(shape).scatter_add_(dim, idx, src)
    if reduce in ('sum', 'add'):
        return res
    if reduce == 'mean':
        ones = torch.ones(index.shape, dtype=src.dtype, device=src.device)
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).scatter_add_(0, index, ones)
        cnt = cnt.clamp(min=1)
        view = [1] * src.dim()
        view[dim] = dim_size
        return res / cnt.view(view)
    raise NotImplementedError(reduce)


def scatter_softmax(*a, **k):
    raise NotImplementedError('scatter_softmax: not on the YOLaT hot path')
