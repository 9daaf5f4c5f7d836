"""Generate tests/golden/convfamily.pt from the UNMODIFIED reference (build container only):

    python oracle/make_golden_family.py

The sibling GraphConv recipes of the hot-path conv -- 'edge', 'attr_edge', 'multilayer_edge', 'attr_edge_gp'
(/root/reference/gcn_lib/sparse/torch_vertex.py:738-747) -- constructed with act='relu', norm='batch' (what YOLaT's
Backbone passes, architecture3cc_rpn_gp_iter2.py:19-20), run through oracle/shims in fp64 (anchor) and fp32 (forward
only) on small random graphs with duplicate edges, a self loop and a node without incoming edges; explicit upstream
gradients; BatchNorm buffers after one training step; eval-mode outputs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader          # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'convfamily.pt')


def fixture(gl, conv, N, E, Cin, C, seed, with_weight):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    c0 = gl.GraphConv(Cin, C, conv, 'relu', 'batch', True)
    for p in c0.parameters():
        p.data = torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.5)
    sd = {k: v.clone() for k, v in c0.state_dict().items()}
    xw = 2 * Cin if conv == 'attr_edge_gp' else Cin          # attr_edge_gp reads [features | root features]
    x = torch.randn(N, xw, generator=g)
    edge = torch.randint(0, N, (E, 2), generator=g)
    edge[1] = edge[0]
    edge[2, 1] = edge[2, 0]
    edge[:, 1][edge[:, 1] == N - 1] = 0
    attr = torch.randn(E, 4, generator=g)
    w = torch.rand(E, generator=g) if with_weight else None
    go = torch.randn(N, C, generator=g)
    res = {}
    for dt in (torch.float64, torch.float32):
        c = gl.GraphConv(Cin, C, conv, 'relu', 'batch', True)
        c.load_state_dict(sd)
        c = c.to(dt).train()
        xx = x.to(dt).requires_grad_(True)
        out = c(xx, edge.t(), None if w is None else w.to(dt), None if conv == 'edge' else attr.to(dt))
        (out * go.to(dt)).sum().backward()
        tag = '64' if dt == torch.float64 else '32'
        res['out' + tag] = out.detach()
        after = {k: v.clone() for k, v in c.state_dict().items() if 'running' in k or 'num_batches' in k}
        c.eval()
        eo = c(xx.detach(), edge.t(), None if w is None else w.to(dt), None if conv == 'edge' else attr.to(dt))
        if dt == torch.float64:
            res['dx64'] = xx.grad
            res['dparams64'] = {k: (p.grad if p.grad is not None else None) for k, p in c.named_parameters()}
            res['buffers_after64'] = after
            res['eval_out64'] = eo.detach()
    return dict(conv=conv, N=N, E=E, Cin=Cin, C=C, state=sd, x=x, edge=edge, attr=attr, edge_weight=w, grad_out=go, **res)


def main():
    ref_loader.load()
    import gcn_lib.sparse as gl
    cases = []
    seed = 100
    for conv in ('edge', 'attr_edge', 'multilayer_edge', 'attr_edge_gp'):
        for (N, E, Cin, C, ww) in ((40, 150, 64, 64, False), (30, 100, 5, 64, True)):
            seed += 1
            cases.append(fixture(gl, conv, N, E, Cin, C, seed, ww))
            print('  %-16s N=%d E=%d Cin=%d weighted=%s  |out| %.3f' % (conv, N, E, Cin, ww, float(cases[-1]['out64'].abs().max())))
    torch.save(cases, OUT)
    print('wrote', OUT, '%.1f KB' % (os.path.getsize(OUT) / 1024))


if __name__ == '__main__':
    main()
