"""SparseCADGCN.predict: the device-side proposal slicing (csrc/slicing.cu) against the plain-Python restatement of the
reference's loops (oracle/predict_slicing.py; architecture3cc_rpn_gp_iter2.py:153-234), and the two-stage inference
end to end against running the model on oracle-built slices."""
import pytest
import torch

from test_host import _proposal_forest

pytestmark = pytest.mark.gpu


def _nodes(data, slices):
    root_nodes, child_nodes = [], []
    for i in range(len(slices['roots']) - 1):
        for root in data.roots[slices['roots'][i]:slices['roots'][i + 1]]:
            root_nodes.append((root, i))
            child_nodes += [(c, i) for c in root.children]
    return root_nodes, child_nodes


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_device_slicing_matches_the_reference_loops(seed):
    from oracle import predict_slicing as P
    from yolat_vectorgraphicsrecognition_b200 import synth
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    data, slices = _proposal_forest(seed, n_images=3)
    model = arch.SparseCADGCN(synth.make_opt(n_classes=5)).cuda()
    dd = model._device_data(data)
    for nodes in _nodes(data, slices):
        if not nodes:
            continue
        sp, se, sb = P.ranges(nodes, slices)
        ref = P.build_data(data, sp, se, sb)
        got, bbox = model._build_data_device(dd, nodes, slices)
        assert bbox == sb
        for k in ('x', 'pos', 'bbox_idx', 'edge', 'e_attr', 'bbox', 'stat_feats'):
            assert torch.equal(getattr(got, k).cpu(), getattr(ref, k)), (seed, k)


def test_predict_two_stage_equals_forward_on_oracle_slices():
    """predict(): roots first, then the children of the roots classified as 'object' (last class), interleaved per
    image -- checked against forward() on the slices the reference's python loops produce."""
    from oracle import predict_slicing as P
    from yolat_vectorgraphicsrecognition_b200 import synth
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    data, slices = _proposal_forest(5, n_images=3)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(synth.make_opt(n_classes=2)).cuda().eval()
    with torch.no_grad():
        pred_cls, pred_bbox, _, slice_bbox, slice_image_bbox, _ = model.predict(data, slices)
        root_nodes, _ = _nodes(data, slices)
        sp, se, sb = P.ranges(root_nodes, slices)
        ref_root, _ = model.forward(P.build_data(data, sp, se, sb), slices)
        has_object = (ref_root.max(1)[1] == model.n_classes - 1).cpu()
        child_nodes, count = [], 0
        per_image = []
        for i in range(len(slices['roots']) - 1):
            n_root = n_child = 0
            for root in data.roots[slices['roots'][i]:slices['roots'][i + 1]]:
                if has_object[count]:
                    child_nodes += [(c, i) for c in root.children]
                    n_child += len(root.children)
                count += 1
                n_root += 1
            per_image.append((n_root, n_child))
        if child_nodes:
            sp2, se2, sb2 = P.ranges(child_nodes, slices)
            ref_child, _ = model.forward(P.build_data(data, sp2, se2, sb2), slices)
        rows, r0, c0 = [], 0, 0
        for n_root, n_child in per_image:
            rows.append(ref_root[r0:r0 + n_root]); r0 += n_root
            if child_nodes:
                rows.append(ref_child[c0:c0 + n_child]); c0 += n_child
        ref = torch.cat(rows, 0)
    assert pred_cls.shape == ref.shape
    assert float((pred_cls - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    assert slice_image_bbox[-1] == ref.shape[0]


def test_predict_matches_the_unmodified_reference_golden():
    """SparseCADGCN.predict on the device against the outputs of the UNMODIFIED reference's predict (fp64, CPU,
    tests/golden/predict_forest.pt from oracle/make_golden_predict.py): the same proposals in the same order, logits
    within the 1e-4 forward bar, boxes expanded by 5 % (:339-352), per-image slices identical."""
    from types import SimpleNamespace
    from util import load_golden, max_rel, FWD_TOL
    from test_oracle import _forest_from_plain, _predict_case_state
    for case in load_golden('predict_forest.pt'):
        opt, model = _predict_case_state(case)
        model = model.cuda().eval()
        data = SimpleNamespace(**case['data'])
        data.roots = _forest_from_plain(case['forest'])
        with torch.no_grad():
            pred_cls, pred_bbox, _, slice_bbox, slice_image_bbox, _ = model.predict(data, case['slices'])
        assert tuple(pred_cls.shape) == tuple(case['pred_cls64'].shape)
        assert max_rel(pred_cls, case['pred_cls64']) < FWD_TOL
        assert max_rel(pred_bbox, case['pred_bbox64']) < 1e-6
        assert [int(v) for v in torch.as_tensor(slice_bbox).reshape(-1)] == case['slice_bbox']
        assert [int(v) for v in slice_image_bbox] == case['slice_image_bbox']
