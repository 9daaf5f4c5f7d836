"""Batch assembly for the hot path: one pinned host buffer, one host->device copy.

The reference moves six tensors per step with six synchronous pageable copies inside `SparseCADGCN.forward`
(cad_recognition/architecture3cc_rpn_gp_iter2.py:107-115) and a seventh for the labels inside the loss (:371),
after `collate` has concatenated the per-image tensors (train.py:123-171).  `PackedBatch` is what a collate
function writes instead: x | bbox_idx | edge | bbox | e_attr | labels laid out back to back (256-byte aligned
sections) in ONE page-locked buffer, with the same field names and dtypes as the reference's `Data` object, so the
batch reaches the GPU as a single asynchronous copy and `GraphedStep` can stage the next batch on a copy stream
while the current one is being computed.  The fields are views into the pinned buffer: filling them in place
(`PackedBatch.like(...)` then `pb.x.copy_(...)`) costs no extra host copy.
"""
from types import SimpleNamespace

import torch

FIELDS = ('x', 'bbox_idx', 'edge', 'bbox', 'e_attr', 'labels')
_ALIGN = 256


def _layout(specs):
    """specs: [(name, shape, dtype)] -> ([(name, offset, nbytes, shape, dtype)], total bytes)"""
    out, off = [], 0
    for name, shape, dtype in specs:
        nbytes = int(torch.Size(shape).numel()) * torch.empty((), dtype=dtype).element_size()
        out.append((name, off, nbytes, tuple(shape), dtype))
        off += (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    return out, max(off, _ALIGN)


def _views(buf, layout):
    ns = {}
    for name, off, nbytes, shape, dtype in layout:
        ns[name] = buf[off:off + nbytes].view(dtype).view(shape)
    return ns


class PackedBatch(object):
    """The tensors of one training / inference step in a single pinned host buffer (`.host`, uint8)."""

    def __init__(self, specs, pin=True):
        self.layout, self.nbytes = _layout(specs)
        self.host = torch.empty(self.nbytes, dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            self.host = self.host.pin_memory()
        for k, v in _views(self.host, self.layout).items():
            setattr(self, k, v)

    @classmethod
    def from_batch(cls, batch, pin=True):
        """Pack the FIELDS of any object carrying them as CPU tensors (a reference `Data`, synth.GraphBatch ...)."""
        specs = []
        for f in FIELDS:
            t = getattr(batch, f)
            if t.device.type != 'cpu':
                raise ValueError('PackedBatch.from_batch packs host tensors; %s is on %s' % (f, t.device))
            specs.append((f, tuple(t.shape), t.dtype))
        pb = cls(specs, pin=pin)
        for f in FIELDS:
            getattr(pb, f).copy_(getattr(batch, f))
        for k, v in vars(batch).items():          # python-side metadata (graph counts, slices ...) rides along
            if not torch.is_tensor(v) and not hasattr(pb, k):
                setattr(pb, k, v)
        return pb

    def signature(self):
        return tuple((name, shape, dtype) for name, _, _, shape, dtype in self.layout)

    def payload_bytes(self):
        return sum(nbytes for _, _, nbytes, _, _ in self.layout)

    def device_twin(self, device):
        """(device uint8 buffer of the same layout, namespace of typed views into it)"""
        buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        return buf, SimpleNamespace(**_views(buf, self.layout))
