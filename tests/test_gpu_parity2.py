"""Round-2 parity cases ("close the parity holes"): full-size configs against the fp64 oracle restatement,
the > L2 roofline shape against the restatement on the whole tensor, channel counts that take the unfused edge path,
the YOLaT++-shaped union graph (edges cross proposals), the E = 0 training case, and the debugging paths selected by
environment variables (run in sub-processes: the library reads them once).

Tolerances as in test_gpu_ops.py / test_gpu_model.py: forward 1e-4 (max-abs / max|ref|); per-op backward 1e-4
L2-relative vs fp64 given identical upstream gradients; end-to-end gradients against twice the reference's own
fp32-vs-fp64 noise floor at that size (5e-3)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from util import ROOT, max_rel, l2_rel, FWD_TOL

pytestmark = pytest.mark.gpu

BWD_TOL = 1e-4


def _conv_pair(Cin, C, seed):
    """A GraphConv('attr_edge_gp2') with non-trivial parameters on the GPU + its state as an fp64 restatement dict."""
    from yolat_vectorgraphicsrecognition_b200.gcn_lib.sparse import GraphConv
    from oracle import restatement as R
    g = torch.Generator().manual_seed(seed)
    conv = GraphConv(Cin, C, 'attr_edge_gp2')
    for p in conv.parameters():
        p.data = torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.5)
    state = R.clone_state({'c.' + k: v for k, v in conv.state_dict().items()}, torch.float64)
    return conv.cuda().train(), state


def _ref_grads(state, x, xn, edge, attr, ew, go, gn, dtype):
    from oracle import restatement as R
    st = R.clone_state({k: v.detach() for k, v in state.items()}, dtype)
    xr = x.detach().to(dtype).clone().requires_grad_(True)
    xnr = xn.detach().to(dtype).clone().requires_grad_(True)
    out, on = R.gp2_conv(st, 'c.gconv', xr, xnr, edge.t(), attr.to(dtype), True,
                         edge_weight=None if ew is None else ew.to(dtype))
    params = {k: v for k, v in st.items() if v.requires_grad}
    grads = torch.autograd.grad((out * go.to(dtype)).sum() + (on * gn.to(dtype)).sum(),
                                [xr, xnr] + list(params.values()), allow_unused=True)
    return out.detach(), on.detach(), params, grads


def _check_conv(conv, state, x, xn, edge, attr, ew=None, fwd_tol=FWD_TOL, bwd_tol=BWD_TOL, seed=0, noise_floor=False):
    """fwd + bwd of the CUDA conv against the fp64 restatement with the same upstream gradients.

    noise_floor: at millions of post-BN activations a handful sit within rounding distance of the ReLU boundary, so no
    fp32 implementation reproduces the fp64 gradients to 1e-4 ("excluding elements whose
    pre-activation |y| < 1e-5" -- which cannot be excluded from an aggregated gradient after the fact).  One flipped
    mask moves a statistic-like gradient (a BatchNorm bias: a signed sum over ~E/2 edges) by ~1/sqrt(E/2) of one
    channel, i.e. ~1e-3 of the tensor at E = 5e4; with a forward accuracy of 1e-6 the expected number of flips among
    E*C = 3e6 activations is O(3) (tools/bwd_debug2.py: the round-1 tape backward and the recompute backward show the
    SAME errors, so they are the forward's masks, not the backward's arithmetic).  The per-tensor bound is therefore
    max(1e-4, 2 x the fp32 restatement's own error against fp64, 5e-3 * min(1, E*C / 3e6)) -- 5e-3 is the end-to-end
    gradient bound of test_gpu_model.py; small cases (the goldens, E*C < 1e5) keep the plain 1e-4."""
    g = torch.Generator().manual_seed(1000 + seed)
    N, C = x.shape[0], conv.gconv.lin_r.weight.shape[0]
    go, gn = torch.randn(N, C, generator=g), torch.randn(N, C, generator=g)
    out64, on64, params, grads = _ref_grads(state, x, xn, edge, attr, ew, go, gn, torch.float64)
    noise = None
    if noise_floor:
        _, _, _, g32 = _ref_grads(state, x, xn, edge, attr, ew, go, gn, torch.float32)
        noise = [0.0 if a is None else l2_rel(a, b) for a, b in zip(g32, grads)]

    flip_allowance = 5e-3 * min(1.0, edge.shape[0] * C / 3e6)

    def tol(i):
        return bwd_tol if noise is None else max(bwd_tol, 2.0 * noise[i], flip_allowance)
    xc = x.detach().cuda().requires_grad_(True)
    xnc = xn.detach().cuda().requires_grad_(True)
    out, on = conv(xc, edge.t().cuda(), None if ew is None else ew.cuda(), attr.cuda(), x_node=xnc)
    assert max_rel(out, out64) < fwd_tol, ('out', max_rel(out, out64))
    assert max_rel(on, on64) < fwd_tol, ('xnode', max_rel(on, on64))
    ((out * go.cuda()).sum() + (on * gn.cuda()).sum()).backward()
    assert l2_rel(xc.grad, grads[0]) < tol(0), ('dx', l2_rel(xc.grad, grads[0]), tol(0))
    assert l2_rel(xnc.grad, grads[1]) < tol(1), ('dxnode', l2_rel(xnc.grad, grads[1]), tol(1))
    got = dict(conv.named_parameters())
    for i, ((k, _), ref) in enumerate(zip(params.items(), grads[2:])):
        p = got[k[2:]]
        if ref is None or float(ref.abs().max()) < 1e-9 * max(1.0, float(go.abs().max())):
            assert float(p.grad.abs().max()) < 2e-5 * (1 + N / 1000.0), k      # bias feeding a training-mode BN
        else:
            assert l2_rel(p.grad, ref) < tol(2 + i), (k, l2_rel(p.grad, ref), tol(2 + i))


@pytest.mark.parametrize('config', ['floorplans', 'diagrams'])
def test_model_matches_oracle_full_config(config):
    """Configs 2 and 3 of BASELINE.json at FULL size (batch 4) against the fp64 restatement: logits / loss 1e-4,
    gradients within the reference's own noise floor, BN running buffers 1e-4."""
    from yolat_vectorgraphicsrecognition_b200 import synth
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    from oracle import restatement as R
    make, kw = synth.CONFIGS[config]
    opt = synth.make_opt(**kw)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt)
    st = R.clone_state(model.state_dict(), torch.float64)
    batch = make()
    assert batch.x.shape[0] == (20000 if config == 'floorplans' else 12000)
    ref = R.run_step(st, opt, batch, training=True)
    ref32 = R.run_step(R.clone_state(model.state_dict(), torch.float32), opt, batch, training=True)
    model = model.cuda().train()
    out = model(batch, None)
    loss = arch.DetectionLoss(opt)(out, batch)['loss']
    loss.backward()
    assert max_rel(out[0], ref['logits']) < FWD_TOL, max_rel(out[0], ref['logits'])
    assert abs(float(loss.detach()) - float(ref['loss'])) < FWD_TOL
    worst = {}
    for k, p in model.named_parameters():
        g = ref['grads'][k]
        if float(g.abs().max()) < 1e-12:
            assert float(p.grad.abs().max()) < 2e-5, k
        else:
            # End-to-end gradients are not reproducible to 1e-4 by ANY fp32 implementation at this size (SURVEY.md
            # fact 10, section 8c: 2-5e-3 expected at config 2): the per-proposal arg-max of 1.28 M (proposal, channel)
            # pairs and ~25 M ReLU masks flip wherever two candidates are closer than the forward's rounding error,
            # and every flip re-routes one gradient element (tools/e2e_noise.py: the fp32 restatement itself is 1-2e-3
            # off its fp64 run on most tensors; this engine's worst tensor 2.6e-3, median 1.7e-3 with the
            # round-to-nearest 3xTF32 split).  Bound: max(5e-3, 2 x the fp32 restatement's own error on the tensor).
            rel = float((p.grad.double().cpu() - g).norm() / g.norm())
            noise = float((ref32['grads'][k].double() - g).norm() / g.norm())
            bound = max(5e-3, 2.0 * noise)
            worst[k] = (rel, noise)
            assert rel < bound, (k, rel, noise)
    for k, v in st.items():
        if 'running' in k:
            assert max_rel(model.state_dict()[k], v) < FWD_TOL, k


def test_gp2conv_roofline_shape_vs_oracle():
    """The > L2 roofline shape (batch 64 of config 2: N = 320 000, E = 1 280 000): forward AND backward of the block
    conv against the fp64 restatement on the whole tensors (the fused K-EDGE passes in both directions)."""
    from yolat_vectorgraphicsrecognition_b200 import synth
    b = synth.floorplans_batch(graphs=64, seed=7)
    conv, state = _conv_pair(64, 64, 11)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(b.x.shape[0], 64, generator=g)
    xn = torch.randn(b.x.shape[0], 64, generator=g)
    _check_conv(conv, state, x, xn, b.edge, b.e_attr, seed=1, noise_floor=True)


@pytest.mark.parametrize('C', [32, 128])
def test_gp2conv_other_channel_counts(C):
    """n_filters 32 / 128 take the unfused edge path (z1 -> GEMM -> aggregate, tape backward)."""
    g = torch.Generator().manual_seed(C)
    N, E = 700, 2900
    conv, state = _conv_pair(C, C, C)
    edge = torch.randint(0, N, (E, 2), generator=g)
    _check_conv(conv, state, torch.randn(N, C, generator=g), torch.randn(N, C, generator=g), edge,
                torch.randn(E, 4, generator=g), ew=torch.rand(E, generator=g), seed=C)


def test_gp2conv_hierarchical_union_graph():
    """Config 5 (YOLaT++-shaped three-level union graph, 15 000 nodes / 50 000 edges per graph): edges cross
    proposals and node degrees are skewed -- per-op parity of the head (5 -> 64) and block (64 -> 64) convs."""
    from yolat_vectorgraphicsrecognition_b200 import synth
    b = synth.hierarchical_batch(graphs=1)
    N = b.x.shape[0]
    g = torch.Generator().manual_seed(9)
    conv, state = _conv_pair(5, 64, 21)
    _check_conv(conv, state, b.x, b.x, b.edge, b.e_attr, seed=2, noise_floor=True)
    conv, state = _conv_pair(64, 64, 22)
    _check_conv(conv, state, torch.randn(N, 64, generator=g), torch.randn(N, 64, generator=g), b.edge, b.e_attr,
                ew=torch.rand(b.edge.shape[0], generator=g), seed=3, noise_floor=True)


def test_gp2conv_zero_edges_training():
    """E = 0 in training mode.  The reference cannot run this case (nn.BatchNorm1d over zero rows raises
    'Expected more than 1 value per channel'); here the edge branch vanishes: out = lin_r(x), the edge-MLP gradients
    are zero and its BatchNorm buffers stay untouched, the node branch is the usual Linear + BN + ReLU."""
    from oracle import restatement as R
    N = 300
    conv, state = _conv_pair(64, 64, 3)
    g = torch.Generator().manual_seed(4)
    x, xn = torch.randn(N, 64, generator=g), torch.randn(N, 64, generator=g)
    xc = x.cuda().requires_grad_(True)
    before = {k: v.clone() for k, v in conv.state_dict().items()}
    out, on = conv(xc, torch.zeros(2, 0, dtype=torch.long).cuda(), None, torch.zeros(0, 4).cuda(), x_node=xn.cuda())
    ref = torch.nn.functional.linear(x.double(), state['c.gconv.lin_r.weight'], state['c.gconv.lin_r.bias'])
    assert max_rel(out, ref) < FWD_TOL
    assert max_rel(on, R.mlp(state, 'c.gconv.mlp_node', xn.double(), 1, True)) < FWD_TOL
    (out.sum() + on.sum()).backward()
    for k, p in conv.named_parameters():
        if k.startswith('gconv.nn.'):
            assert float(p.grad.abs().max()) == 0.0, k
    sd = conv.state_dict()
    for k in ('gconv.nn.1.running_mean', 'gconv.nn.4.running_var', 'gconv.nn.1.num_batches_tracked'):
        assert torch.equal(sd[k], before[k]), k
    assert int(sd['gconv.mlp_node.1.num_batches_tracked']) == 1


_ENV_SCRIPT = r'''
import json, sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
from util import load_golden, max_rel, l2_rel
from yolat_vectorgraphicsrecognition_b200.gcn_lib.sparse import GraphConv
res = {}
for name in ('head', 'block_weighted'):
    fx = load_golden('gp2conv_%%s.pt' %% name)
    conv = GraphConv(fx['Cin'], fx['C'], 'attr_edge_gp2'); conv.load_state_dict(fx['state']); conv = conv.cuda().train()
    x = fx['x'].detach().cuda().requires_grad_(True); xn = fx['x_node'].detach().cuda().requires_grad_(True)
    w = fx['edge_weight'].cuda() if fx['edge_weight'] is not None else None
    out, on = conv(x, fx['edge'].t().cuda(), w, fx['attr'].cuda(), x_node=xn)
    (out * fx['grad_out'].cuda()).sum().add((on * fx['grad_xnode'].cuda()).sum()).backward()
    worst = max(l2_rel(p.grad, fx['dparams64'][k]) for k, p in conv.named_parameters() if float(fx['dparams64'][k].abs().max()) > 1e-9)
    res[name] = dict(out=max_rel(out, fx['out64']), dx=l2_rel(x.grad, fx['dx64']), dparams=worst)
print('RESULT ' + json.dumps(res))
'''


@pytest.mark.parametrize('env', [{'YOLAT_EDGE_BWD': 'tape'}, {'YOLAT_EDGE': 'unfused'}, {'YOLAT_GEMM': 'simt'},
                                 {'YOLAT_BRANCHES': '0'}])
def test_debug_paths_keep_parity(env):
    """The alternative CUDA paths selected by environment variables (tape backward of round 1, unfused edge path, SIMT
    GEMM, single-stream schedule) must meet the same bars as the default path."""
    script = _ENV_SCRIPT % dict(root=ROOT, tests=os.path.join(ROOT, 'tests'))
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, '-c', script], env=e, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith('RESULT ')][-1][7:])
    for name, r in res.items():
        assert r['out'] < FWD_TOL and r['dx'] < BWD_TOL and r['dparams'] < BWD_TOL, (env, name, r)
