"""Drop-in for the `torch_scatter.scatter` calls on the YOLaT hot path
(cad_recognition/architecture3cc_rpn_gp_iter2.py:8,67,122): dim 0, 1-D index, reduce mean / max / sum.

Semantics restated from torch-scatter 2.0.x (third-party, not under /root/reference): mean divides by
clamp(count, 1); empty segments give 0; max routes its gradient to a single arg index (first
occurrence).  Runs on the segmented (atomic-free) sm_100a kernels.
"""
import torch

from . import ops
from .graph import Segments


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    if out is not None:
        raise NotImplementedError('yolat_b200 scatter: `out=` is not supported')
    if dim < 0:
        dim = src.dim() + dim
    if dim != 0 or (not isinstance(index, Segments) and index.dim() != 1):
        raise NotImplementedError('yolat_b200 scatter: only dim=0 with a 1-D index is on the YOLaT hot path')
    seg = index if isinstance(index, Segments) else Segments(index, dim_size)
    shape = src.shape
    flat = src.reshape(shape[0], -1)
    if reduce == 'max':
        res = ops.segment_max(flat, seg)
    elif reduce == 'mean':
        res = ops.segment_mean(flat, seg)
    else:
        raise NotImplementedError('yolat_b200 scatter: reduce=%r is not on the YOLaT hot path' % (reduce,))
    return res.reshape((seg.S,) + tuple(shape[1:]))
