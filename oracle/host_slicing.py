"""TEST INFRASTRUCTURE (oracle): vectorised HOST mirror of predict()'s range lists and `build_data`
(cad_recognition/architecture3cc_rpn_gp_iter2.py:153-234 of the reference), used by tests/test_host.py to pin the
slicing logic against the plain-Python restatement (oracle/predict_slicing.py) without a GPU.  The product path never
imports this file: SparseCADGCN.predict slices on the device (csrc/slicing.cu)."""
import numpy as np
import torch


def ranges(nodes, slices, key_lists):
    """Concatenated index ranges for a list of (tree node, image index) pairs."""
    out = {k: [] for k in ('pos', 'edge', 'edge_super')}
    bbox = []
    for node, i in nodes:
        v = node.value
        out['pos'].append(np.arange(v['idx_pos'][0], v['idx_pos'][1]) + int(slices['pos'][i]))
        out['edge'].append(np.arange(v['idx_edge'][0], v['idx_edge'][1]) + int(slices['edge'][i]))
        bbox.append(int(v['idx_bbox'] + slices['bbox'][i]))
    cat = {k: (np.concatenate(v) if len(v) else np.zeros(0, dtype=np.int64)).astype(np.int64) for k, v in out.items()
           if k != 'edge_super'}
    return cat['pos'], cat['edge'], bbox

def build_data(data, slice_pos, slice_edge, slice_bbox):
    """build_data (:167-234) without the per-edge / per-node python loops: old->new node renumbering is
    a lookup table, the dense bbox_idx is a run-length cumsum."""
    from types import SimpleNamespace
    sp = torch.as_tensor(slice_pos, dtype=torch.long)
    se = torch.as_tensor(slice_edge, dtype=torch.long)
    x_all = data.x
    o2n = torch.full((x_all.shape[0],), -1, dtype=torch.long)
    o2n[sp] = torch.arange(sp.numel())
    nd = SimpleNamespace()
    nd.x = data.x.cpu()[sp] if not data.x.is_cuda else data.x[sp.to(data.x.device)]
    nd.pos = data.pos[sp] if getattr(data, 'pos', None) is not None else None
    old_bidx = data.bbox_idx.cpu()[sp]
    nd.edge = o2n[data.edge.cpu()[se]]
    nd.e_attr = data.e_attr.cpu()[se]
    sb = torch.as_tensor(slice_bbox, dtype=torch.long)
    nd.bbox = data.bbox.cpu()[sb]
    nd.stat_feats = data.stat_feats.cpu()[sb] if getattr(data, 'stat_feats', None) is not None else None
    if old_bidx.numel() > 0:
        change = torch.zeros_like(old_bidx)
        change[1:] = (old_bidx[1:] != old_bidx[:-1]).long()
        nd.bbox_idx = torch.cumsum(change, 0)
    else:
        nd.bbox_idx = old_bidx
    return nd

