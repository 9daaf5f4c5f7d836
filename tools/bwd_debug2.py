"""Debug aid: block-wise errors of dW1 (W1a | W1b | W1c columns) and the other gradients of a GraphConv on the config-5
union graph, against the fp64 restatement and its own fp32 run.  YOLAT_EDGE_BWD=tape python tools/bwd_debug2.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
from util import l2_rel, max_rel  # noqa: E402
import test_gpu_parity2 as T  # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import synth  # noqa: E402


def run(use_ew, which):
    if which == 'hier':
        b = synth.hierarchical_batch(graphs=1)
    else:
        b = synth.floorplans_batch(graphs=3, seed=3)
    N = b.x.shape[0]
    g = torch.Generator().manual_seed(9)
    conv, state = T._conv_pair(64, 64, 22)
    x, xn = torch.randn(N, 64, generator=g), torch.randn(N, 64, generator=g)
    ew = torch.rand(b.edge.shape[0], generator=g) if use_ew else None
    gg = torch.Generator().manual_seed(1003)
    go, gn = torch.randn(N, 64, generator=gg), torch.randn(N, 64, generator=gg)
    o64, n64, params, g64 = T._ref_grads(state, x, xn, b.edge, b.e_attr, ew, go, gn, torch.float64)
    _, _, _, g32 = T._ref_grads(state, x, xn, b.edge, b.e_attr, ew, go, gn, torch.float32)
    xc = x.cuda().requires_grad_(True)
    xnc = xn.cuda().requires_grad_(True)
    out, on = conv(xc, b.edge.t().cuda(), None if ew is None else ew.cuda(), b.e_attr.cuda(), x_node=xnc)
    ((out * go.cuda()).sum() + (on * gn.cuda()).sum()).backward()
    print('--- %s ew=%s  mode=%s  N=%d E=%d  fwd %.2e' % (which, use_ew, os.environ.get('YOLAT_EDGE_BWD', 'recompute'), N,
                                                       b.edge.shape[0], max_rel(out, o64)))
    print('   dx %.2e (ref32 %.2e)' % (l2_rel(xc.grad, g64[0]), l2_rel(g32[0], g64[0])))
    got = dict(conv.named_parameters())
    for i, (k, _) in enumerate(params.items()):
        r64, r32, p = g64[2 + i], g32[2 + i], got[k[2:]]
        if r64 is None or float(r64.abs().max()) < 1e-9:
            continue
        line = '   %-24s %.2e (ref32 %.2e)' % (k[2:], l2_rel(p.grad, r64), l2_rel(r32, r64))
        if k.endswith('nn.0.weight'):
            for name, sl in (('W1a', slice(0, 64)), ('W1b', slice(64, 128)), ('W1c', slice(128, 132))):
                line += '  %s %.2e (%.2e)' % (name, l2_rel(p.grad[:, sl], r64[:, sl]), l2_rel(r32[:, sl], r64[:, sl]))
        print(line)


if __name__ == '__main__':
    for which in ('hier', 'floor'):
        for use_ew in (False, True):
            run(use_ew, which)
