"""oracle/predict_slicing.py -- TEST INFRASTRUCTURE.  Plain-Python restatement of the host-side slicing inside
`SparseCADGCN.predict` (cad_recognition/architecture3cc_rpn_gp_iter2.py): the root / child range lists
(:153-162, :276-290) and `build_data` (:167-234: old->new node dictionary, per-edge re-indexing loop, per-node
bbox_idx renumbering loop).  Loops are kept as the reference writes them; only the PyG `Data` container is replaced
by a namespace.  Used by tests/test_host.py to check the vectorised host mirror."""
from types import SimpleNamespace

import numpy as np
import torch


def ranges(nodes, slices):
    """nodes: [(tree node, image index)] in visiting order -> (slice_pos, slice_edge, slice_bbox)  (:153-162)"""
    slice_pos, slice_edge, slice_bbox = [], [], []
    for node, i in nodes:
        v = node.value
        slice_pos += list(range(v['idx_pos'][0] + int(slices['pos'][i]), v['idx_pos'][1] + int(slices['pos'][i])))
        slice_edge += list(range(v['idx_edge'][0] + int(slices['edge'][i]), v['idx_edge'][1] + int(slices['edge'][i])))
        slice_bbox.append(int(v['idx_bbox'] + slices['bbox'][i]))
    return slice_pos, slice_edge, slice_bbox


def build_data(data, slice_pos, slice_edge, slice_bbox):
    """(:167-234)"""
    o2n = {}
    for new_i, old_i in enumerate(slice_pos):
        o2n[old_i] = new_i
    nd = SimpleNamespace(x=data.x[slice_pos], pos=data.pos[slice_pos])
    nd.bbox_idx = data.bbox_idx[slice_pos]
    edge = []
    for e in data.edge[slice_edge].numpy():
        edge.append([o2n[e[0]], o2n[e[1]]])
    nd.edge = torch.tensor(edge, dtype=torch.long).reshape(-1, 2)
    nd.e_attr = data.e_attr[slice_edge]
    nd.bbox = data.bbox[slice_bbox]
    nd.stat_feats = data.stat_feats[slice_bbox]
    new_bbox_idx = [0]
    count = 0
    for i in range(1, nd.bbox_idx.size(0)):
        if nd.bbox_idx[i] != nd.bbox_idx[i - 1]:
            count += 1
        new_bbox_idx.append(count)
    nd.bbox_idx = torch.tensor(np.array(new_bbox_idx), dtype=torch.long)
    return nd
