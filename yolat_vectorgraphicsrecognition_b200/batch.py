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


OFFSET_FIELD = 'offset_tab'      # [4][G+1] int64: edge | pos | bbox_idx | labels slices (train.py:238-258)


def collate(data_list, pin=True):
    """Drop-in for `train.collate` (cad_recognition/train.py:123-171): same `(data, slices)` contract -- every key of
    the per-image `Data` objects is concatenated along its cat dim, python lists are joined, `slices[key]` are the
    prefix sums as int64 tensors -- but the tensors the model reads (FIELDS) are written straight into ONE pinned
    buffer (no intermediate torch.cat copies), together with the four slice tables the per-image offset loop of
    train.py:238-258 needs.  `data` is a PackedBatch:

      * handed to the reference's own training loop it behaves like the reference's collated `Data`: the host loop adds
        the offsets in place (the fields are views of the pinned buffer) and `model(data, slices)` uploads one buffer;
      * `data.defer_offsets()` marks the offsets to be applied on the DEVICE instead (csrc/slicing.cu
        `yolat_batch_offsets`, one launch, captured in the step graph): skip the host loop.
    """
    first = data_list[0]
    keys = list(first.keys)
    G = len(data_list)
    slices = {k: [0] for k in keys}
    sizes = {}
    for item in data_list:
        for k in keys:
            v = item[k]
            if torch.is_tensor(v) and v.dim() > 0:
                cd = item.__cat_dim__(k, v)
                cd = 0 if cd is None else cd
                n = v.size(cd)
            elif isinstance(v, list):
                n = len(v)
            else:
                n = 1
            slices[k].append(slices[k][-1] + n)
    specs = []
    for f in FIELDS:
        v = first[f]
        if not torch.is_tensor(v):
            raise ValueError('collate: field %s must be a tensor' % f)
        cd = first.__cat_dim__(f, v)
        if cd not in (None, 0):
            raise ValueError('collate: field %s is concatenated along dim %s (expected 0)' % (f, cd))
        specs.append((f, (slices[f][-1],) + tuple(v.shape[1:]), v.dtype))
    specs.append((OFFSET_FIELD, (4, G + 1), torch.int64))
    pb = PackedBatch(specs, pin=pin)
    for f in FIELDS:
        dst = getattr(pb, f)
        for i, item in enumerate(data_list):
            dst[slices[f][i]:slices[f][i + 1]].copy_(item[f])
    tab = getattr(pb, OFFSET_FIELD)
    for row, k in enumerate(('edge', 'pos' if 'pos' in slices else 'x', 'bbox_idx', 'labels')):
        tab[row].copy_(torch.tensor(slices[k], dtype=torch.int64))
    pb.offsets_pending = False
    # every other key exactly as the reference collate treats it
    for k in keys:
        if k in FIELDS:
            continue
        vals = [item[k] for item in data_list]
        v0 = vals[0]
        if torch.is_tensor(v0) and G > 1:
            if v0.dim() > 0:
                cd = first.__cat_dim__(k, v0)
                setattr(pb, k, torch.cat(vals, dim=0 if cd is None else cd))
            else:
                setattr(pb, k, torch.stack(vals))
        elif torch.is_tensor(v0):
            setattr(pb, k, v0)
        elif isinstance(v0, (int, float)):
            setattr(pb, k, torch.tensor(vals))
        elif isinstance(v0, list):
            out = []
            for v in vals:
                out += v
            setattr(pb, k, out)
        else:
            setattr(pb, k, vals)
    pb._keys = keys
    return pb, {k: torch.tensor(v, dtype=torch.long) for k, v in slices.items()}


def _pb_keys(self):
    return list(getattr(self, '_keys', FIELDS))


def _pb_defer_offsets(self):
    """The per-image offsets will be added on the device (apply_offsets): do not run the host loop of train.py:238-258."""
    if not hasattr(self, OFFSET_FIELD):
        raise ValueError('defer_offsets needs a batch built by batch.collate')
    self.offsets_pending = True
    return self


def apply_offsets(ns, stream=None):
    """edge / bbox_idx of the DEVICE namespace `ns` (a device twin of a collated PackedBatch) += per-image offsets."""
    from . import _lib as L
    tab = getattr(ns, OFFSET_FIELD)
    G = tab.shape[1] - 1
    L.check(L.lib().yolat_batch_offsets(ns.edge.data_ptr(), ns.edge.shape[0], ns.bbox_idx.data_ptr(), ns.bbox_idx.shape[0],
                                        tab.data_ptr(), G, L.stream() if stream is None else stream), 'yolat_batch_offsets')


PackedBatch.keys = property(_pb_keys)
PackedBatch.__getitem__ = lambda self, k: getattr(self, k, None)
PackedBatch.__setitem__ = lambda self, k, v: setattr(self, k, v)
PackedBatch.defer_offsets = _pb_defer_offsets
