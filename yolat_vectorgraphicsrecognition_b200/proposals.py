"""Proposal enumeration on the device -- drop-in for `SESYDFloorPlan._get_proposal`.

Reference: `Datasets/graph_dict3.py:309-789`.  The reference enumerates, per connected component of the Bezier graph,
every node set cut out by windows of a `bbox_sampling_step` grid, filters / labels / measures each set in python
loops (seconds per image; cached to `*_bb.pkl` afterwards, `:922-932`).  Here the whole function is one C-ABI call
pair (`yolat_proposals_count` -> sizes, `yolat_proposals_fill` -> outputs; csrc/proposals.cu) around one pinned H2D
copy of the image's graph.  There is no CPU path: without the CUDA library or a CUDA device this raises.

Same arguments, same 14-tuple, same exceptions as the reference (`ValueError` for a component without x or y extent,
`SystemExit` for a component no ground-truth box touches, `ValueError` where `np.argmax([])` / `np.concatenate([])`
would raise, `KeyError` for an edge that names a control point).  One documented difference: the proposals of one
component come out in first-occurrence order of the window walk, where the reference's order is CPython's `set`
iteration order (`list(set(sub_clusters))`, `:557`).
`do_mixup` (random augmentation, `:791-903`) is not part of this path.
"""
import ctypes as C
import gc

import numpy as np
import torch

from . import _lib

N_STATS = 13


class idxTree(object):
    """graph_dict3.py:24-27."""

    def __init__(self):
        self.children = []
        self.value = {}


class ProposalIn(C.Structure):
    _fields_ = [('pos', C.c_void_p), ('is_control', C.c_void_p), ('is_super', C.c_void_p), ('n_all', C.c_int64),
                ('cc_ptr', C.c_void_p), ('cc_idx', C.c_void_p), ('ncc', C.c_int64), ('cc_total', C.c_int64),
                ('edge', C.c_void_p), ('e_attr', C.c_void_p), ('E', C.c_int64), ('A', C.c_int32),
                ('edge_super', C.c_void_p), ('e_attr_super', C.c_void_p), ('Es', C.c_int64), ('As', C.c_int32),
                ('gt_bbox', C.c_void_p), ('gt_labels', C.c_void_p), ('G', C.c_int64),
                ('sampling_step', C.c_int32), ('n_classes', C.c_int32), ('normalize_bbox', C.c_int32)]


class ProposalOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'pos', 'is_super', 'bbox_idx', 'edge', 'edge_super', 'e_attr', 'e_attr_super', 'labels', 'has_obj', 'bbox',
        'bbox_targets', 'stat_feats', 'slice_pos', 'slice_edge', 'slice_super', 'slice_bbox', 'cc_table')]


# totals[] slots and error bits of include/yolat_b200.h
T_NODES, T_EDGES, T_SUPER, T_BOXES, T_ERR, T_ERR_CC, T_NODES_IN, N_TOTALS = 0, 1, 2, 3, 4, 5, 6, 8
ERR_ZERO_STEP, ERR_NO_GT, ERR_NO_PROPOSAL, ERR_CONTROL_REF, ERR_CC_OVERLAP, ERR_LIMIT, ERR_INDEX = 1, 2, 4, 8, 16, 32, 64

# (name, dtype, columns) of the packed input sections, in buffer order
_SECTIONS = ('pos', 'e_attr', 'e_attr_super', 'gt_bbox', 'cc_ptr', 'cc_idx', 'edge', 'edge_super', 'gt_labels',
             'is_control', 'is_super')


def pack_graph_dict(graph_dict, gt_bbox, gt_labels):
    """The reference's inputs (graph_dict3.py:310-319) as flat numpy arrays + sizes.  Host-side marshalling only: no
    arithmetic on the data.  `cc` (list of lists) becomes CSR; 8-byte sections first so one buffer holds them aligned."""
    pos = np.ascontiguousarray(np.asarray(graph_dict['pos']['spatial'], dtype=np.float64).reshape(-1, 2))
    n_all = pos.shape[0]
    is_control_src = np.asarray(graph_dict['attr']['is_control'])
    is_super_src = np.asarray(graph_dict['attr']['is_super'])
    cc = graph_dict['cc']
    cc_ptr = np.zeros(len(cc) + 1, dtype=np.int64)
    if len(cc):
        cc_ptr[1:] = np.cumsum([len(c) for c in cc])
    cc_idx = (np.concatenate([np.asarray(c, dtype=np.int64).reshape(-1) for c in cc]) if len(cc) and cc_ptr[-1] > 0
              else np.zeros(0, dtype=np.int64))

    def edges(key):
        e = np.asarray(graph_dict['edge'][key], dtype=np.int64).reshape(-1, 2)
        a = np.asarray(graph_dict['edge_attr'][key], dtype=np.float64)
        a = a.reshape(e.shape[0], -1) if e.shape[0] else np.zeros((0, a.shape[1] if a.ndim == 2 and a.shape[1] else 1))
        if a.shape[0] != e.shape[0]:
            raise ValueError('edge_attr[%r] has %d rows for %d edges' % (key, a.shape[0], e.shape[0]))
        return np.ascontiguousarray(e), np.ascontiguousarray(a)

    edge, e_attr = edges('shape')
    edge_super, e_attr_super = edges('super')
    gt_bbox = np.ascontiguousarray(np.asarray(gt_bbox, dtype=np.float64).reshape(-1, 4))
    gt_labels = np.ascontiguousarray(np.asarray(gt_labels, dtype=np.int64).reshape(-1))
    if gt_labels.shape[0] != gt_bbox.shape[0]:
        raise ValueError('gt_labels and gt_bbox disagree')
    if is_control_src.reshape(-1).shape[0] != n_all or is_super_src.reshape(-1).shape[0] != n_all:
        raise ValueError('is_control / is_super must have one entry per node')
    return {
        'pos': pos, 'e_attr': e_attr, 'e_attr_super': e_attr_super, 'gt_bbox': gt_bbox,
        'cc_ptr': cc_ptr, 'cc_idx': np.ascontiguousarray(cc_idx), 'edge': edge, 'edge_super': edge_super,
        'gt_labels': gt_labels,
        'is_control': np.ascontiguousarray((is_control_src.reshape(-1) != 0).astype(np.uint8)),
        'is_super': np.ascontiguousarray((is_super_src.reshape(-1) != 0).astype(np.uint8)),
        'is_super_dtype': is_super_src.dtype,
    }


def fill_in_struct(p, addr, sampling_step, n_classes, normalize_bbox):
    """ProposalIn over packed arrays; `addr(name)` gives the (device) address of section `name`."""
    s = ProposalIn()
    for name in _SECTIONS:
        setattr(s, name, addr(name) if p[name].size else None)
    s.cc_ptr = addr('cc_ptr')
    s.n_all = p['pos'].shape[0]
    s.ncc = p['cc_ptr'].shape[0] - 1
    s.cc_total = p['cc_idx'].shape[0]
    s.E, s.A = p['edge'].shape[0], max(p['e_attr'].shape[1], 1)
    s.Es, s.As = p['edge_super'].shape[0], max(p['e_attr_super'].shape[1], 1)
    s.G = p['gt_bbox'].shape[0]
    s.sampling_step, s.n_classes, s.normalize_bbox = int(sampling_step), int(n_classes), int(bool(normalize_bbox))
    return s


def output_specs(totals, A, As, ncc):
    """(name, numpy dtype, shape) of every output, sized from the totals of yolat_proposals_count."""
    n, e, es, b = (int(totals[k]) for k in (T_NODES, T_EDGES, T_SUPER, T_BOXES))
    return [('pos', np.float64, (n, 2)), ('e_attr', np.float64, (e, A)), ('e_attr_super', np.float64, (es, As)),
            ('bbox', np.float64, (b, 4)), ('bbox_targets', np.float64, (b, 4)), ('stat_feats', np.float64, (b, N_STATS)),
            ('bbox_idx', np.int64, (n,)), ('edge', np.int64, (e, 2)), ('edge_super', np.int64, (es, 2)),
            ('labels', np.int64, (b,)), ('has_obj', np.int64, (b,)),
            ('slice_pos', np.int64, (b + 1,)), ('slice_edge', np.int64, (b + 1,)), ('slice_super', np.int64, (b + 1,)),
            ('slice_bbox', np.int64, (b + 1,)), ('cc_table', np.int64, (ncc, 3)), ('is_super', np.uint8, (n,))]


def raise_for(totals):
    """Map the device error mask to the exception the reference raises at the same condition."""
    err, cc = int(totals[T_ERR]), int(totals[T_ERR_CC])
    if not err:
        return
    if err & ERR_INDEX:
        raise IndexError('proposal enumeration: a node id lies outside [0, n_all)')
    if err & ERR_CONTROL_REF:
        raise KeyError('proposal enumeration: an edge or a component names a control point (o2n lookup, graph_dict3.py:332-348)')
    if err & ERR_CC_OVERLAP:
        raise ValueError('proposal enumeration: component %d shares a node with another component' % cc)
    if err & ERR_LIMIT:
        raise ValueError('proposal enumeration: component %d exceeds the kernel limits (2^20-2 nodes, 2^23-1 edges, '
                         '62 grid cells)' % cc)
    if err & ERR_ZERO_STEP:
        raise ValueError('arange: cannot compute length (component %d has no x or y extent: np.arange with step 0, '
                         'graph_dict3.py:470-476)' % cc)
    if err & ERR_NO_GT:
        print('cc has no intersect gt bbox')          # graph_dict3.py:575-576
        raise SystemExit
    if err & ERR_NO_PROPOSAL:
        raise ValueError('attempt to get argmax of an empty sequence (component %d yields no proposal, '
                         'graph_dict3.py:728)' % cc)
    raise _lib.YolatError('proposal enumeration: error mask %#x' % err)


def unpack_outputs(o, is_super_dtype):
    """The reference's 14-tuple (graph_dict3.py:787) from the flat outputs."""
    if o['edge_super'].shape[0] == 0 or o['edge'].shape[0] == 0:
        raise ValueError('need at least one array to concatenate')      # np.concatenate([]) at :772-773
    # the idxTree objects are the reference's return type (:730-750); python lists + dict literals keep their
    # construction at ~1 us per proposal
    sp, se, ss, sb = (o[k].tolist() for k in ('slice_pos', 'slice_edge', 'slice_super', 'slice_bbox'))
    pos_r, edge_r, sup_r = list(zip(sp[:-1], sp[1:])), list(zip(se[:-1], se[1:])), list(zip(ss[:-1], ss[1:]))
    new = idxTree.__new__

    def node(i):
        t = new(idxTree)
        t.children = []
        t.value = {'idx_pos': pos_r[i], 'idx_edge': edge_r[i], 'idx_edge_super': sup_r[i], 'idx_bbox': sb[i]}
        return t

    roots = []
    gc_on = gc.isenabled()
    gc.disable()                  # thousands of small containers: keep the cyclic collector from rescanning the heap
    try:
        for first, count, root_i in o['cc_table'].tolist():
            root = node(root_i)
            root.children = [node(i) for i in range(first, first + count) if i != root_i]
            roots.append(root)
    finally:
        if gc_on:
            gc.enable()
    pos = o['pos']
    return (pos, o['is_super'].reshape(-1, 1).astype(is_super_dtype), np.zeros((pos.shape[0], 1)), o['edge'],
            o['edge_super'], o['e_attr'], o['e_attr_super'], o['labels'].tolist(), o['bbox_idx'], o['bbox'],
            o['bbox_targets'], o['stat_feats'], o['has_obj'].tolist(), roots)


def _sections_layout(sizes):
    """256-byte aligned offsets of (name, nbytes) sections in one buffer."""
    offs, off = {}, 0
    for name, nbytes in sizes:
        offs[name] = off
        off += (nbytes + 255) // 256 * 256
    return offs, max(off, 256)


def get_proposal(graph_dict, gt_bbox, gt_labels, bbox_sampling_step=5, n_classes=17, normalize_bbox=True,
                 device=None):
    """`SESYDFloorPlan._get_proposal(graph_dict, gt_bbox, gt_labels, bbox_sampling_step)` on the GPU.

    Returns (pos, is_super, is_control, edge, edge_super, e_attr, e_attr_super, labels, bbox_idx, bbox, bbox_targets,
    stat_feats, has_obj, roots) exactly as graph_dict3.py:787 does (numpy arrays / python lists / idxTree roots)."""
    lib = _lib.lib()
    if not torch.cuda.is_available():
        raise _lib.YolatError('yolat_b200 has no CPU path: proposal enumeration needs a CUDA device')
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    p = pack_graph_dict(graph_dict, gt_bbox, gt_labels)
    with torch.cuda.device(device):
        # one pinned buffer, one H2D copy
        arrays = [(name, p[name]) for name in _SECTIONS]
        offs, nbytes = _sections_layout([(name, a.nbytes) for name, a in arrays])
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        hv = host.numpy()
        for name, a in arrays:
            hv[offs[name]:offs[name] + a.nbytes] = a.reshape(-1).view(np.uint8)
        dev = host.to(device, non_blocking=True)
        base = dev.data_ptr()
        s_in = fill_in_struct(p, lambda name: base + offs[name], bbox_sampling_step, n_classes, normalize_bbox)
        ws_bytes = lib.yolat_proposals_ws_bytes(C.byref(s_in))
        if ws_bytes < 0:
            raise ValueError('proposal enumeration: unsupported sizes (bbox_sampling_step must be in 1..62)')
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        totals_d = torch.empty(N_TOTALS, dtype=torch.int64, device=device)
        st = _lib.stream()
        _lib.check(lib.yolat_proposals_count(C.byref(s_in), ws.data_ptr(), ws_bytes, totals_d.data_ptr(), st),
                   'proposals_count')
        totals = totals_d.cpu().numpy()               # the one synchronisation: output sizes
        raise_for(totals)
        specs = output_specs(totals, s_in.A, s_in.As, s_in.ncc)
        nbytes_of = lambda dt, shape: int(np.prod(shape, dtype=np.int64)) * np.dtype(dt).itemsize
        ooffs, onbytes = _sections_layout([(name, nbytes_of(dt, shape)) for name, dt, shape in specs])
        out_d = torch.empty(onbytes, dtype=torch.uint8, device=device)
        s_out = ProposalOut()
        for name, _, _ in specs:
            setattr(s_out, name, out_d.data_ptr() + ooffs[name])
        _lib.check(lib.yolat_proposals_fill(C.byref(s_in), ws.data_ptr(), ws_bytes, C.byref(s_out), st), 'proposals_fill')
        out_h = torch.empty(onbytes, dtype=torch.uint8, pin_memory=True)
        out_h.copy_(out_d, non_blocking=True)         # one D2H copy of all outputs
        torch.cuda.current_stream().synchronize()
        ov = out_h.numpy()
        o = {}
        for name, dt, shape in specs:
            o[name] = ov[ooffs[name]:ooffs[name] + nbytes_of(dt, shape)].view(dt).reshape(shape).copy()
    return unpack_outputs(o, p['is_super_dtype'])


class ProposalEnumerator(object):
    """Holds the three Dataset attributes `_get_proposal` reads (`n_classes`, `normalize_bbox`, `do_mixup`;
    graph_dict3.py:53-54,104) so that `SESYDFloorPlan._get_proposal = ProposalEnumerator._get_proposal` style
    overlays keep the reference's call `self._get_proposal(graph_dict, gt_bbox, gt_labels, bbox_sampling_step=...)`."""

    def __init__(self, n_classes=17, normalize_bbox=True, do_mixup=False):
        self.n_classes, self.normalize_bbox, self.do_mixup = n_classes, normalize_bbox, do_mixup

    def _get_proposal(self, graph_dict, gt_bbox, gt_labels, bbox_sampling_step=-1):
        if getattr(self, 'do_mixup', False):
            raise NotImplementedError('do_mixup (random augmentation, graph_dict3.py:791-903) is not on this path')
        if bbox_sampling_step < 1:
            raise ValueError('bbox_sampling_step must be positive (the reference divides the box by it)')
        return get_proposal(graph_dict, gt_bbox, gt_labels, bbox_sampling_step, self.n_classes, self.normalize_bbox)
