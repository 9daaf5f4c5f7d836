"""Seeded synthetic Bezier-graph batches in the layout the reference Dataset emits.

No dataset files ship with the reference (only list files), so every BASELINE.json config is
generated.  Layout follows `Datasets/graph_dict3.py:966-971` + `train.collate`
(`cad_recognition/train.py:123-171`) + the per-image offset loop (`train.py:238-258`), i.e. the
tensors `SparseCADGCN.forward` reads (`architecture3cc_rpn_gp_iter2.py:107-115`):

    x        [N,5]  fp32   = [0,0,0,pos_x,pos_y]
    edge     [E,2]  int64  (src j, dst i) -- one direction only, already offset into the batch
    e_attr   [E,4]  fp32   control-point offsets [c0-start, c1-end]
    bbox_idx [N]    int64  proposal id of each node, non-decreasing, contiguous
    bbox     [B,4]  fp32   echoed through as pred_bbox
    stat_feats [B,13] fp32 copied to the device by the reference but unused (dim_stat = 0)
    labels   [B]    int64
    is_super [N]    int64  read (and ignored) by DetectionLoss

The definitions are the ones SURVEY.md section 8(d) fixes for configs 1-5.
"""
import math

import torch


class GraphBatch(object):
    """Minimal attribute bag with the PyG-1.x `Data` surface the reference touches."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith('_')]

    def __getitem__(self, k):
        return getattr(self, k, None)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def __contains__(self, k):
        return k in self.keys

    def __cat_dim__(self, key, value):
        return 0

    def tensors(self):
        return {k: v for k, v in self.__dict__.items() if torch.is_tensor(v)}

    def to(self, device, non_blocking=False):
        out = GraphBatch(**self.__dict__)
        for k, v in self.tensors().items():
            setattr(out, k, v.to(device, non_blocking=non_blocking))
        return out

    def pin_memory(self):
        out = GraphBatch(**self.__dict__)
        for k, v in self.tensors().items():
            setattr(out, k, v.pin_memory())
        return out

    def input_bytes(self):
        """Bytes of the tensors forward + loss move host->device (x, bbox_idx, edge, bbox, stat_feats, e_attr, labels)."""
        names = ('x', 'bbox_idx', 'edge', 'bbox', 'stat_feats', 'e_attr', 'labels')
        return sum(getattr(self, n).numel() * getattr(self, n).element_size() for n in names)


def _finish(pos, edge, e_attr, bbox_idx, labels, n_graphs, g):
    N = pos.shape[0]
    B = int(bbox_idx[-1]) + 1 if N > 0 else 0
    x = torch.cat([torch.zeros(N, 3), pos], dim=1).contiguous()
    return GraphBatch(
        x=x, pos=pos, edge=edge.contiguous(), e_attr=e_attr.contiguous(), bbox_idx=bbox_idx,
        bbox=torch.rand(B, 4, generator=g), stat_feats=torch.zeros(B, 13), labels=labels,
        is_super=torch.zeros(N, dtype=torch.long), n_graphs=n_graphs)


def _edges_inside_proposals(seg_start, seg_len, E, g):
    """Per edge: owner proposal ~ U{0..B-1}; src,dst ~ U{0..len-1} inside it (self loops / duplicates allowed)."""
    B = seg_len.numel()
    owner = torch.randint(0, B, (E,), generator=g)
    ln = seg_len[owner]
    src = (torch.rand(E, generator=g, dtype=torch.float64) * ln).long().clamp_(max=ln.max() - 1)
    dst = (torch.rand(E, generator=g, dtype=torch.float64) * ln).long().clamp_(max=ln.max() - 1)
    src = torch.minimum(src, ln - 1) + seg_start[owner]
    dst = torch.minimum(dst, ln - 1) + seg_start[owner]
    return torch.stack([src, dst], dim=1)


def floorplans_batch(graphs=4, n=5000, e=20000, npp=16, n_classes=17, seed=1):
    """Config 2 (and, with graphs=64, the > L2 scale-up point): fixed 16-node proposals."""
    g = torch.Generator().manual_seed(seed)
    N, E = graphs * n, graphs * e
    bbox_idx = torch.arange(N) // npp
    B = int(bbox_idx[-1]) + 1
    seg_start = torch.arange(B) * npp
    seg_len = torch.clamp(N - seg_start, max=npp)
    pos = torch.rand(N, 2, generator=g)
    edge = _edges_inside_proposals(seg_start, seg_len, E, g)
    e_attr = torch.randn(E, 4, generator=g) * 0.1
    labels = torch.randint(0, n_classes, (B,), generator=g)
    return _finish(pos, edge, e_attr, bbox_idx, labels, graphs, g)


def diagrams_batch(graphs=4, n=3000, e=9000, n_classes=22, seed=2, lo=3, hi=12):
    """Config 3: variable-length proposals npp ~ U{3..12} (bbox_sampling_step=5 => more, smaller proposals)."""
    g = torch.Generator().manual_seed(seed)
    N, E = graphs * n, graphs * e
    lens = torch.randint(lo, hi + 1, (N // lo + 1,), generator=g)
    cs = torch.cumsum(lens, 0)
    B = int((cs < N).sum()) + 1
    lens = lens[:B].clone()
    lens[B - 1] = N - (int(cs[B - 2]) if B > 1 else 0)
    seg_start = torch.cumsum(lens, 0) - lens
    bbox_idx = torch.repeat_interleave(torch.arange(B), lens)
    pos = torch.rand(N, 2, generator=g)
    edge = _edges_inside_proposals(seg_start, lens, E, g)
    e_attr = torch.randn(E, 4, generator=g) * 0.1
    labels = torch.randint(0, n_classes, (B,), generator=g)
    return _finish(pos, edge, e_attr, bbox_idx, labels, graphs, g)


_KAPPA = 0.552284749831  # circle -> 4 cubic Beziers, Datasets/bezier_parser.py:103-131


def toy_batch(n_nodes=128, n_shapes=32, seed=0):
    """Config 1: restatement of the `toy_dataset.py:57-83` shape generator without svgpathtools.

    circle -> 4 cubic Beziers (kappa = 0.5523), rectangle -> 4 lines, triangle -> 3 lines; nodes are
    the curve end points (control points dropped, toy_dataset.py:138-150); one proposal per shape;
    e_attr = control-point offsets (zeros for lines: a line is the degenerate cubic whose controls
    sit on its end points, bezier_parser.py:65-70); padded with rectangles to exactly `n_nodes`.
    Classes: circle 0, triangle 1, rectangle 2 (toy_dataset.py:24-30).
    """
    g = torch.Generator().manual_seed(seed)

    def u(*shape):
        return torch.rand(*shape, generator=g, dtype=torch.float64)

    pos, edge, attr, bidx, labels = [], [], [], [], []
    kinds = torch.randint(0, 3, (n_shapes,), generator=g).tolist()
    budget = n_nodes
    shapes = []
    for k in kinds:
        need = 3 if k == 1 else 4
        if budget - need < 0:
            break
        shapes.append(k)
        budget -= need
    while budget >= 4:
        shapes.append(2)
        budget -= 4
    if budget == 3:
        shapes.append(1)
        budget = 0
    assert budget == 0, 'cannot pad to exactly n_nodes'
    base = 0
    for b, k in enumerate(shapes):
        if k == 0:  # circle
            r = float(u(1)) * 0.5
            cx = cy = 0.5
            m = r * _KAPPA
            p = [(cx, cy - r), (cx + r, cy), (cx, cy + r), (cx - r, cy)]
            c = [((cx + m, cy - r), (cx + r, cy - m)), ((cx + r, cy + m), (cx + m, cy + r)),
                 ((cx - m, cy + r), (cx - r, cy + m)), ((cx - r, cy - m), (cx - m, cy - r))]
            for q in range(4):
                s, t = q, (q + 1) % 4
                edge.append(tuple(sorted((base + s, base + t))))
                attr.append([c[q][0][0] - p[s][0], c[q][0][1] - p[s][1], c[q][1][0] - p[t][0], c[q][1][1] - p[t][1]])
            labels.append(0)
        elif k == 2:  # rectangle
            w, h = float(u(1)), float(u(1))
            p = [(0.0, 0.0), (w, 0.0), (w, h), (0.0, h)]
            for q in range(4):
                edge.append(tuple(sorted((base + q, base + (q + 1) % 4))))
                attr.append([0.0] * 4)
            labels.append(2)
        else:  # triangle
            xy = u(3, 2).tolist()
            p = [tuple(v) for v in xy]
            for q in range(3):
                edge.append(tuple(sorted((base + q, base + (q + 1) % 3))))
                attr.append([0.0] * 4)
            labels.append(1)
        pos += p
        bidx += [b] * len(p)
        base += len(p)
    pos = torch.tensor(pos, dtype=torch.float32)
    return _finish(pos, torch.tensor(edge, dtype=torch.long), torch.tensor(attr, dtype=torch.float32),
                   torch.tensor(bidx, dtype=torch.long), torch.tensor(labels, dtype=torch.long), 1, g)


def hierarchical_batch(graphs=1, n_p=12000, n_c=2500, n_r=500, seed=5, n_classes=17):
    """Config 5 (shape only; no YOLaT++ code exists in the reference): three-level union graph.

    points / curves / primitives; edges 30000 p-p, 12000 p->c, 5000 c-c, 2500 c->r, 500 r-r per graph;
    bbox_idx = primitive id (points and curves are assigned to primitives in contiguous runs).
    """
    g = torch.Generator().manual_seed(seed)
    pos_all, edge_all, bidx_all = [], [], []
    nb = 0
    nprop = 0
    for _ in range(graphs):
        n = n_p + n_c + n_r
        # contiguous proposal ids: order nodes by primitive
        prim_of_p = torch.sort(torch.randint(0, n_r, (n_p,), generator=g)).values
        prim_of_c = torch.sort(torch.randint(0, n_r, (n_c,), generator=g)).values
        prim_of_r = torch.arange(n_r)
        prim = torch.cat([prim_of_p, prim_of_c, prim_of_r])
        order = torch.sort(prim, stable=True).indices
        inv = torch.empty_like(order)
        inv[order] = torch.arange(n)
        P0, C0, R0 = 0, n_p, n_p + n_c

        def r(lo, hi, m):
            return torch.randint(lo, hi, (m,), generator=g)
        e = torch.cat([
            torch.stack([r(P0, C0, 30000), r(P0, C0, 30000)], 1),
            torch.stack([r(P0, C0, 12000), r(C0, R0, 12000)], 1),
            torch.stack([r(C0, R0, 5000), r(C0, R0, 5000)], 1),
            torch.stack([r(C0, R0, 2500), r(R0, n, 2500)], 1),
            torch.stack([r(R0, n, 500), r(R0, n, 500)], 1)])
        edge_all.append(inv[e] + nb)
        pos_all.append(torch.rand(n, 2, generator=g))
        bidx_all.append(prim[order] + nprop)
        nb += n
        nprop += n_r
    pos = torch.cat(pos_all)
    edge = torch.cat(edge_all)
    bbox_idx = torch.cat(bidx_all)
    e_attr = torch.randn(edge.shape[0], 4, generator=g) * 0.1
    labels = torch.randint(0, n_classes, (nprop,), generator=g)
    return _finish(pos, edge, e_attr, bbox_idx, labels, graphs, g)


def make_opt(n_classes=17, in_channels=5, n_blocks=2, n_blocks_out=2, n_filters=64, **kw):
    """The `opt` fields the model reads (config.py:24-86; README training commands)."""
    from types import SimpleNamespace
    d = dict(n_filters=n_filters, act='relu', norm='batch', bias=True, conv='attr_edge',
             n_blocks=n_blocks, n_blocks_out=n_blocks_out, n_classes=n_classes, class_specific=False,
             in_channels=in_channels, classifier='softmax', dropout=0.0)
    d.update(kw)
    return SimpleNamespace(**d)


CONFIGS = {
    # name: (batch factory, opt kwargs)
    'toy': (lambda **k: toy_batch(**k), dict(n_classes=3, n_blocks=1, n_blocks_out=1)),
    'floorplans': (lambda **k: floorplans_batch(**k), dict(n_classes=17)),
    'diagrams': (lambda **k: diagrams_batch(**k), dict(n_classes=22)),
    'hierarchical': (lambda **k: hierarchical_batch(**k), dict(n_classes=17)),
}
