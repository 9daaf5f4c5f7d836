"""GPU parity of the sibling GraphConv recipes ('edge', 'attr_edge', 'multilayer_edge', 'attr_edge_gp')
against golden vectors of the UNMODIFIED reference (tests/golden/convfamily.pt, oracle/make_golden_family.py):
forward 1e-4, backward 1e-4 vs fp64 with identical upstream gradients, BatchNorm buffers, eval mode, state-dict layout."""
import pytest
import torch

from util import load_golden, max_rel, l2_rel, FWD_TOL

pytestmark = pytest.mark.gpu

CASES = load_golden('convfamily.pt')


@pytest.mark.parametrize('i', range(len(CASES)), ids=['%s-%d' % (c['conv'], c['Cin']) for c in CASES])
def test_conv_family_matches_reference(i):
    from yolat_vectorgraphicsrecognition_b200.gcn_lib.sparse import GraphConv
    fx = CASES[i]
    conv = GraphConv(fx['Cin'], fx['C'], fx['conv'], 'relu', 'batch', True)
    assert list(conv.state_dict()) == list(fx['state'])            # same children, same order, same names
    conv.load_state_dict(fx['state'])
    conv = conv.cuda().train()
    x = fx['x'].detach().cuda().requires_grad_(True)
    w = fx['edge_weight'].cuda() if fx['edge_weight'] is not None else None
    attr = None if fx['conv'] == 'edge' else fx['attr'].cuda()
    out = conv(x, fx['edge'].t().cuda(), w, attr)
    assert max_rel(out, fx['out64']) < FWD_TOL, max_rel(out, fx['out64'])
    (out * fx['grad_out'].cuda()).sum().backward()
    assert l2_rel(x.grad, fx['dx64']) < 1e-4, l2_rel(x.grad, fx['dx64'])
    for k, p in conv.named_parameters():
        ref = fx['dparams64'][k]
        if ref is None:                               # children the reference constructs but never uses (mlp, lin_l)
            assert p.grad is None, k
        elif float(ref.abs().max()) < 1e-9:          # bias feeding a training-mode BN
            assert float(p.grad.abs().max()) < 2e-5, k
        else:
            assert l2_rel(p.grad, ref) < 1e-4, (k, l2_rel(p.grad, ref))
    sd = conv.state_dict()
    for k, v in fx['buffers_after64'].items():
        if 'num_batches' in k:
            assert int(sd[k]) == int(v), k
        else:
            assert max_rel(sd[k], v) < FWD_TOL, k
    conv.eval()
    with torch.no_grad():
        eo = conv(x.detach(), fx['edge'].t().cuda(), w, attr)
    assert max_rel(eo, fx['eval_out64']) < FWD_TOL


def test_conv_family_rejects_other_settings():
    from yolat_vectorgraphicsrecognition_b200.gcn_lib.sparse import GraphConv
    with pytest.raises(NotImplementedError):
        GraphConv(8, 64, 'attr_edge', 'relu', None, True)          # norm=None is not built
    with pytest.raises(NotImplementedError):
        GraphConv(8, 64, 'gat')
    with pytest.raises(NotImplementedError, match='conv nope is not implemented'):
        GraphConv(8, 64, 'nope')
