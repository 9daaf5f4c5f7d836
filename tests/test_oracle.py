"""CPU: the oracle restatement (oracle/restatement.py) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/*.pt, written by oracle/make_golden.py in the build container)."""
from types import SimpleNamespace

import pytest
import torch

from util import load_golden, l2_rel, max_rel, sample
from oracle import restatement as R
from yolat_vectorgraphicsrecognition_b200 import synth
from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch

CASES = {
    'toy': lambda: synth.toy_batch(),
    'floorplans_small': lambda: synth.floorplans_batch(graphs=2, n=400, e=1600, seed=1),
    'diagrams_small': lambda: synth.diagrams_batch(graphs=2, n=300, e=900, seed=2),
    'floorplans_3blk': lambda: synth.floorplans_batch(graphs=1, n=320, e=1200, seed=3),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_restatement_matches_reference_model(name):
    fx = load_golden('model_%s.pt' % name)
    opt = SimpleNamespace(**fx['opt'])
    torch.manual_seed(fx['seed'])
    init = arch.SparseCADGCN(opt).state_dict()       # same seeded construction order as the reference
    for k, v in init.items():
        assert abs(float(v.double().sum()) - fx['init_checksum'][k][0]) < 1e-9, k
        assert abs(float(v.double().abs().sum()) - fx['init_checksum'][k][1]) < 1e-9, k
    batch = CASES[name]()
    st = R.clone_state(init, torch.float64)
    res = R.run_step(st, opt, batch, training=True)
    assert max_rel(res['logits'], fx['logits64']) < 1e-11
    assert abs(float(res['loss']) - fx['loss64']) < 1e-11
    for k, g in res['grads'].items():
        ref = fx['grad64_sample'][k]
        if fx['grad64_absmax'][k] < 1e-12:
            assert float(g.abs().max()) < 1e-12, k
        else:
            assert l2_rel(sample(g), ref) < 1e-8, k
    for k, v in fx['bn_after64'].items():
        assert max_rel(sample(st[k]), v) < 1e-12, k
    # fp32 restatement vs fp32 reference: bit-for-bit the same op chain
    st32 = R.clone_state(init, torch.float32)
    res32 = R.run_step(st32, opt, batch, training=True)
    assert max_rel(res32['logits'], fx['logits32']) < 2e-5
    # eval mode after the step
    with torch.no_grad():
        ev = R.cadgcn_forward(st32, opt, batch.x, batch.edge, batch.e_attr, batch.bbox_idx, training=False)
    assert max_rel(ev, fx['eval_logits32']) < 2e-5


@pytest.mark.parametrize('name', ['head', 'block', 'block_weighted', 'block_sparse'])
def test_restatement_matches_reference_conv(name):
    fx = load_golden('gp2conv_%s.pt' % name)
    st = {('c.' + k).replace('c.gconv', 'c'): (v.double().clone() if v.is_floating_point() else v.clone())
          for k, v in fx['state'].items()}
    for k, v in st.items():
        if v.is_floating_point() and 'running' not in k:
            v.requires_grad_(True)
    x = fx['x'].detach().double().requires_grad_(True)
    xn = fx['x_node'].detach().double().requires_grad_(True)
    w = fx['edge_weight'].double() if fx['edge_weight'] is not None else None
    out, on = R.gp2_conv(st, 'c', x, xn, fx['edge'].t(), fx['attr'].double(), True, w)
    assert max_rel(out, fx['out64']) < 1e-11 and max_rel(on, fx['xnode64']) < 1e-11
    (out * fx['grad_out'].double()).sum().add((on * fx['grad_xnode'].double()).sum()).backward()
    assert l2_rel(x.grad, fx['dx64']) < 1e-9 and l2_rel(xn.grad, fx['dxnode64']) < 1e-9
    for k, ref in fx['dparams64'].items():
        g = st[('c.' + k).replace('c.gconv', 'c')].grad
        if float(ref.abs().max()) < 1e-9:
            assert float(g.abs().max()) < 1e-9
        else:
            assert l2_rel(g, ref) < 1e-9, k
    for k, v in fx['buffers_after64'].items():
        assert max_rel(st[('c.' + k).replace('c.gconv', 'c')], v) < 1e-12


def test_restatement_scatter_matches_reference():
    for case in load_golden('scatter.pt'):
        S = int(case['index'].max()) + 1
        for red, fn in (('mean', R.scatter_mean), ('max', R.scatter_max)):
            s = case['src'].clone().requires_grad_(True)
            o = fn(s, case['index'], S)
            (o * case['grad']).sum().backward()
            assert torch.equal(o, case[red]) and torch.equal(s.grad, case['d' + red])


def test_reference_itself_when_present():
    """In the build container the unmodified reference is importable through oracle/shims: one live check."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('/root/reference is not on this machine')
    try:
        archref = ref_loader.load()
        opt = synth.make_opt(n_classes=17)
        torch.manual_seed(0)
        model = archref.SparseCADGCN(opt).train()
        batch = synth.floorplans_batch(graphs=1, n=160, e=600, seed=8)
        out = model(batch, None)
        st = R.clone_state(model.state_dict(), torch.float32)
        # the forward above already updated the BN buffers once; compare in eval mode on both sides
        model.eval()
        with torch.no_grad():
            a = model(batch, None)[0]
            b = R.cadgcn_forward(st, opt, batch.x, batch.edge, batch.e_attr, batch.bbox_idx, training=False)
        assert max_rel(b, a) < 1e-5 and out[0].shape == (10, 17)
    finally:
        ref_loader.unload()


def _forest_from_plain(plain):
    from types import SimpleNamespace

    def node(d):
        v = {k: (tuple(x) if isinstance(x, list) else x) for k, x in d['value'].items()}
        return SimpleNamespace(value=v, children=[node(c) for c in d['children']])
    return [node(r) for r in plain]


def _predict_case_state(case):
    """The case's weights: same seed + same construction order as the reference => bit-identical init (checked)."""
    import torch
    from yolat_vectorgraphicsrecognition_b200 import synth
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    opt = synth.make_opt(n_classes=case['n_classes'])
    torch.manual_seed(case['seed'])
    model = arch.SparseCADGCN(opt)
    sd = model.state_dict()
    for k, v in case['state_checksum'].items():
        assert abs(float(sd[k].double().sum()) - v) <= 1e-6 * max(1.0, abs(v)), k
    return opt, model


def test_predict_restatement_matches_reference_golden():
    """oracle/predict_slicing.py + oracle/restatement.py against the UNMODIFIED reference's SparseCADGCN.predict
    (tests/golden/predict_forest.pt, generated by oracle/make_golden_predict.py): same proposals classified, in the
    same order, with the same logits -- this pins the slicing restatement the GPU tests are checked against."""
    import torch
    from types import SimpleNamespace
    from oracle import predict_slicing as P, restatement as R
    for case in load_golden('predict_forest.pt'):
        opt, model = _predict_case_state(case)
        state = R.clone_state(model.state_dict(), torch.float64, requires_grad=False)
        data = SimpleNamespace(**{k: (v.double() if v.is_floating_point() else v) for k, v in case['data'].items()})
        data.roots = _forest_from_plain(case['forest'])
        slices = case['slices']
        n_img = len(slices['roots']) - 1
        root_nodes = [(r, i) for i in range(n_img) for r in data.roots[slices['roots'][i]:slices['roots'][i + 1]]]
        sp, se, sb = P.ranges(root_nodes, slices)
        nd = P.build_data(data, sp, se, sb)
        with torch.no_grad():
            root_cls = R.cadgcn_forward(state, opt, nd.x, nd.edge, nd.e_attr, nd.bbox_idx, training=False)
        has_object = root_cls.max(1)[1] == opt.n_classes - 1
        child_nodes, per_image, count = [], [], 0
        for i in range(n_img):
            n_root = n_child = 0
            for root in data.roots[slices['roots'][i]:slices['roots'][i + 1]]:
                if has_object[count]:
                    child_nodes += [(c, i) for c in root.children]
                    n_child += len(root.children)
                count += 1
                n_root += 1
            per_image.append((n_root, n_child))
        assert child_nodes, 'fixture should exercise the second stage'
        sp2, se2, sb2 = P.ranges(child_nodes, slices)
        nd2 = P.build_data(data, sp2, se2, sb2)
        with torch.no_grad():
            child_cls = R.cadgcn_forward(state, opt, nd2.x, nd2.edge, nd2.e_attr, nd2.bbox_idx, training=False)
        rows, boxes, r0, c0 = [], [], 0, 0
        for n_root, n_child in per_image:
            rows += [root_cls[r0:r0 + n_root], child_cls[c0:c0 + n_child]]
            boxes += sb[r0:r0 + n_root] + sb2[c0:c0 + n_child]
            r0 += n_root; c0 += n_child
        got = torch.cat(rows, 0)
        assert got.shape == case['pred_cls64'].shape
        assert float((got - case['pred_cls64']).abs().max()) < 1e-9
        assert boxes == case['slice_bbox']
