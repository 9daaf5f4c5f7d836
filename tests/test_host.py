"""CPU: host-side mirror of the reference interface -- module tree, state-dict keys, construction-order
RNG parity, stage grouping, error behaviour, synthetic batch layouts, sharding helper."""
from types import SimpleNamespace

import pytest
import torch

from oracle import restatement as R
from yolat_vectorgraphicsrecognition_b200 import synth, dp
from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
from yolat_vectorgraphicsrecognition_b200.gcn_lib import sparse as gl


def test_state_dict_keys_and_shapes_match_reference_layout():
    for kw in (dict(n_classes=17), dict(n_classes=22, n_blocks=3, n_blocks_out=2), dict(n_classes=3, n_blocks=1, n_blocks_out=1)):
        opt = synth.make_opt(**kw)
        sd = arch.SparseCADGCN(opt).state_dict()
        spec = R.state_spec(opt)
        assert [k for k, _, _ in spec] == list(sd.keys())
        for k, shape, _ in spec:
            assert tuple(sd[k].shape) == tuple(shape), k
    assert sum(p.numel() for p in arch.SparseCADGCN(synth.make_opt(n_classes=17)).parameters()) == 1613329


def test_attribute_paths():
    m = arch.SparseCADGCN(synth.make_opt(n_classes=17))
    assert isinstance(m.cls_net.head.gconv.nn, gl.MLP) and isinstance(m.cls_net.backbone[0].body.gconv.lin_r, torch.nn.Linear)
    assert len(m.cls_net.heads) == 0 and m.cls_net.fusion_dims == 128 and m.dim_stat == 0
    assert [type(c).__name__ for c in m.cls_net.head.gconv.nn] == ['Linear', 'BatchNorm1d', 'ReLU'] * 2
    assert [type(c).__name__ for c in m.prediction_cls[2]] == ['Linear']


def test_mlp_constructor_semantics():
    assert [type(c).__name__ for c in gl.MLP([4, 8, 3], 'relu', 'batch', last_lin=True)] == ['Linear', 'BatchNorm1d', 'ReLU', 'Linear']
    assert [type(c).__name__ for c in gl.MLP([4, 8], 'leakyrelu', 'layer', drop=0.5)] == ['Linear', 'LayerNorm', 'LeakyReLU', 'Dropout2d']
    assert gl.MLP([4, 8], None, 'none', bias=False)[0].bias is None
    st = gl.MLP([4, 8, 3], 'relu', 'batch', last_lin=True).stages()
    assert [s[0] for s in st] == ['stage', 'stage'] and st[0][2] is not None and st[0][3] and st[1][2] is None
    with pytest.raises(NotImplementedError):
        gl.act_layer('gelu')
    with pytest.raises(NotImplementedError):
        gl.norm_layer('group', 8)


def test_graphconv_dispatch_errors():
    with pytest.raises(NotImplementedError, match='conv foo is not implemented'):
        gl.GraphConv(4, 8, 'foo')
    with pytest.raises(NotImplementedError):
        gl.GraphConv(4, 8, 'edge')
    with pytest.raises(NotImplementedError):
        gl.PlainDynBlock(8)
    c = gl.GraphConv(5, 64, 'ATTR_EDGE_GP2', act='prelu', norm='layer', bias=False)   # act/norm/bias ignored (:749)
    assert c.conv == 'attr_edge_gp2' and c.gconv.lin_r.bias is not None


def test_multiseq_splats_tuples():
    class Two(torch.nn.Module):
        def forward(self, a, b=None):
            return (a + 1, a * 2) if b is None else a + b
    assert float(gl.MultiSeq(Two(), Two())(torch.tensor(1.0))) == 4.0


def test_no_cpu_path():
    from yolat_vectorgraphicsrecognition_b200 import _lib
    with pytest.raises(_lib.YolatError):
        gl.MLP([8, 8], 'relu', 'batch')(torch.randn(4, 8))
    with pytest.raises(_lib.YolatError):
        from yolat_vectorgraphicsrecognition_b200.torch_scatter import scatter
        scatter(torch.randn(4, 2), torch.tensor([0, 0, 1, 1]), dim=0, reduce='max')


def test_synthetic_layouts():
    b = synth.floorplans_batch()
    assert b.x.shape == (20000, 5) and b.edge.shape == (80000, 2) and b.e_attr.shape == (80000, 4)
    assert b.bbox.shape == (1250, 4) and b.labels.shape == (1250,) and b.edge.dtype == torch.int64
    assert bool((b.bbox_idx[1:] >= b.bbox_idx[:-1]).all()) and int(b.bbox_idx[-1]) == 1249
    assert bool((b.bbox_idx[b.edge[:, 0]] == b.bbox_idx[b.edge[:, 1]]).all())      # edges never cross proposals
    assert bool((b.x[:, :3] == 0).all())
    t = synth.toy_batch()
    assert t.x.shape == (128, 5) and int(t.labels.max()) <= 2
    d = synth.diagrams_batch(graphs=1, n=300, e=900)
    cnt = torch.bincount(d.bbox_idx)
    assert int(cnt.min()) >= 1 and d.bbox.shape[0] == cnt.numel()
    h = synth.hierarchical_batch()
    assert h.x.shape[0] == 15000 and h.edge.shape[0] == 50000 and int(h.bbox_idx[-1]) == 499
    assert torch.equal(synth.floorplans_batch(graphs=1, seed=5).edge, synth.floorplans_batch(graphs=1, seed=5).edge)


def test_shard_graphs():
    for n, w in ((32, 8), (7, 4), (4, 1), (3, 4)):
        parts = [dp.shard_graphs(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    edges = [10, 10, 10, 70, 10, 10, 10, 10]
    parts = [dp.shard_graphs(8, 2, r, edges) for r in range(2)]
    assert parts == [(0, 4), (4, 8)]


def test_packed_batch_layout_roundtrip():
    """batch.PackedBatch: six fields in one buffer, 256-byte aligned sections, views alias the buffer."""
    import torch
    from yolat_vectorgraphicsrecognition_b200 import synth
    from yolat_vectorgraphicsrecognition_b200.batch import FIELDS, PackedBatch
    b = synth.floorplans_batch(graphs=2, n=100, e=300, seed=3)
    pb = PackedBatch.from_batch(b, pin=False)
    assert pb.host.dtype == torch.uint8 and pb.nbytes == pb.host.numel()
    for name, off, nbytes, shape, dtype in pb.layout:
        assert off % 256 == 0
        v = getattr(pb, name)
        assert v.dtype == getattr(b, name).dtype and tuple(v.shape) == tuple(getattr(b, name).shape)
        assert torch.equal(v, getattr(b, name))
        assert v.data_ptr() == pb.host.data_ptr() + off           # a view, not a copy
    assert [l[0] for l in pb.layout] == list(FIELDS)
    assert pb.payload_bytes() <= pb.nbytes
    pb.x.zero_()
    assert float(pb.host[:pb.layout[0][2]].view(torch.float32).abs().sum()) == 0.0
    assert pb.signature() == PackedBatch.from_batch(synth.floorplans_batch(graphs=2, n=100, e=300, seed=9), pin=False).signature()


def _proposal_forest(seed, n_images=2):
    """Synthetic batch with the idxTree structure SparseCADGCN.predict walks (graph_dict3.py:750-768): per image a few
    root proposals, each owning a contiguous node / edge range and a bbox row, with children owning sub-ranges."""
    import torch
    from types import SimpleNamespace
    g = torch.Generator().manual_seed(seed)
    roots, slices = [], {'roots': [0], 'pos': [], 'edge': [], 'bbox': []}
    xs, edges, attrs, bidx, n_bbox = [], [], [], [], 0
    pos_off = edge_off = 0
    for img in range(n_images):
        slices['pos'].append(pos_off); slices['edge'].append(edge_off); slices['bbox'].append(n_bbox)
        p = e = b = 0
        for r in range(3 + img):
            n_r = int(torch.randint(4, 9, (1,), generator=g))
            kids, first_p, first_e = [], p, e
            for c in range(int(torch.randint(0, 3, (1,), generator=g))):
                n_c = int(torch.randint(2, 5, (1,), generator=g))
                e_c = n_c + 1
                kids.append(SimpleNamespace(value={'idx_pos': (p, p + n_c), 'idx_edge': (e, e + e_c), 'idx_bbox': b}, children=[]))
                for _ in range(e_c):
                    edges.append([pos_off + p + int(torch.randint(0, n_c, (1,), generator=g)),
                                  pos_off + p + int(torch.randint(0, n_c, (1,), generator=g))])
                bidx += [n_bbox + b] * n_c
                p += n_c; e += e_c; b += 1
            # the root's own nodes follow its children's and its edges stay inside [first_p, p + n_r)
            for _ in range(n_r):
                edges.append([pos_off + first_p + int(torch.randint(0, p + n_r - first_p, (1,), generator=g)),
                              pos_off + first_p + int(torch.randint(0, p + n_r - first_p, (1,), generator=g))])
            bidx += [n_bbox + b] * n_r
            roots.append(SimpleNamespace(value={'idx_pos': (first_p, p + n_r), 'idx_edge': (first_e, e + n_r), 'idx_bbox': b},
                                         children=kids))
            p += n_r; e += n_r; b += 1
        pos_off += p; edge_off += e; n_bbox += b
        slices['roots'].append(len(roots))
    N, E = pos_off, edge_off
    data = SimpleNamespace(x=torch.randn(N, 5, generator=g), pos=torch.rand(N, 2, generator=g),
                           bbox_idx=torch.tensor(bidx), edge=torch.tensor(edges, dtype=torch.long),
                           e_attr=torch.randn(E, 4, generator=g), bbox=torch.rand(n_bbox, 4, generator=g),
                           stat_feats=torch.rand(n_bbox, 13, generator=g), roots=roots)
    assert data.edge.shape[0] == E and len(bidx) == N
    return data, slices


def test_predict_slicing_matches_the_reference_loops():
    """The vectorised host mirror (oracle/host_slicing.py) against the plain-Python restatement of the
    reference's loops (oracle/predict_slicing.py; architecture3cc_rpn_gp_iter2.py:153-234), on roots and on children."""
    import torch
    from oracle import predict_slicing as P
    from oracle import host_slicing as HS
    for seed in (0, 1, 2):
        data, slices = _proposal_forest(seed)
        root_nodes, child_nodes = [], []
        for i in range(len(slices['roots']) - 1):
            for root in data.roots[slices['roots'][i]:slices['roots'][i + 1]]:
                root_nodes.append((root, i))
                child_nodes += [(c, i) for c in root.children]
        for nodes in (root_nodes, child_nodes):
            if not nodes:
                continue
            sp_ref, se_ref, sb_ref = P.ranges(nodes, slices)
            ref = P.build_data(data, sp_ref, se_ref, sb_ref)
            sp, se, sb = HS.ranges(nodes, slices, None)
            assert list(sp) == sp_ref and list(se) == se_ref and list(sb) == sb_ref
            got = HS.build_data(data, sp, se, sb)
            for k in ('x', 'pos', 'bbox_idx', 'edge', 'e_attr', 'bbox', 'stat_feats'):
                assert torch.equal(getattr(got, k), getattr(ref, k)), (seed, k)


def test_fused_adam_has_no_cpu_path():
    """optim.FusedAdam is CUDA-only like every other op: CPU parameters are rejected loudly, hyper-parameters are
    validated like torch.optim.Adam's."""
    import pytest
    import torch
    from yolat_vectorgraphicsrecognition_b200 import _lib
    from yolat_vectorgraphicsrecognition_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.randn(4, 4))
    opt = FusedAdam([p], lr=1e-3, weight_decay=1e-4)
    assert opt.param_groups[0]['lr'] == 1e-3 and opt.param_groups[0]['betas'] == (0.9, 0.999)
    p.grad = torch.zeros_like(p)
    with pytest.raises(_lib.YolatError):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdam([p], betas=(1.0, 0.999))


def test_f32rows_keeps_column_slices():
    """`_lib.f32rows`: what the backward of torch.cat(dim=1) hands to its producers (column slices of a wider matrix)
    passes through with its pitch; anything else falls back to a contiguous fp32 copy."""
    import torch
    from yolat_vectorgraphicsrecognition_b200 import _lib as L
    wide = torch.arange(60, dtype=torch.float32).reshape(5, 12)
    sl = wide[:, 4:8]
    assert L.f32rows(sl).data_ptr() == sl.data_ptr() and L.f32rows(sl).stride(0) == 12
    assert L.f32rows(wide).data_ptr() == wide.data_ptr()
    t = wide.t()                                         # column-major view: must be copied
    assert L.f32rows(t).is_contiguous() and torch.equal(L.f32rows(t), t)
    e = torch.ones(1, 1).expand(5, 4)                    # stride-0 broadcast (sum().backward()): must be copied
    assert L.f32rows(e).is_contiguous()
    assert L.f32rows(sl.double()).dtype == torch.float32
