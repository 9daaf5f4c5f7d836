"""Generate tests/golden/*.pt from the UNMODIFIED reference (build container only).

    python oracle/make_golden.py

Imports /root/reference/cad_recognition/architecture3cc_rpn_gp_iter2.py and gcn_lib through the
stand-ins in oracle/shims, runs it on seeded synthetic batches in fp32 and fp64, verifies that
oracle/restatement.py reproduces it, and stores small fixtures:

  model_<cfg>.pt   init checksums, logits / loss / grads (sub-sampled for big tensors) / BN buffers after
                   one training step, eval-mode logits, and the reference's own fp32-vs-fp64 noise floor
  gp2conv_<case>.pt  one GraphConv('attr_edge_gp2') with explicit weights + inputs + upstream grads
  scatter.pt       torch_scatter mean/max forward + backward cases (incl. empty segments, ties)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader, restatement as R          # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import synth    # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
SAMPLE_CAP = 1024


def sample(t):
    """Full tensor if small, else an evenly strided sample (indices are reproducible from numel)."""
    t = t.detach().reshape(-1)
    if t.numel() <= SAMPLE_CAP:
        return t.clone()
    step = t.numel() // SAMPLE_CAP
    return t[::step][:SAMPLE_CAP].clone()


def checksum(sd):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def ref_step(arch, opt, batch, dtype, sd=None, training=True, seed=0):
    torch.manual_seed(seed)
    model = arch.SparseCADGCN(opt)
    if sd is not None:
        model.load_state_dict(sd)
    model = model.to(dtype)
    crit = arch.DetectionLoss(opt)
    model.train(training)
    b = synth.GraphBatch(**batch.__dict__)
    b.x = batch.x.to(dtype)
    b.e_attr = batch.e_attr.to(dtype)
    out = model(b, None)
    loss = crit(out, b)['loss']
    if training:
        loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters()} if training else {}
    return model, out[0].detach(), loss.detach(), grads


def model_fixture(arch, name, batch, opt):
    torch.manual_seed(0)
    init = {k: v.clone() for k, v in arch.SparseCADGCN(opt).state_dict().items()}
    m32, lg32, ls32, g32 = ref_step(arch, opt, batch, torch.float32)
    m64, lg64, ls64, g64 = ref_step(arch, opt, batch, torch.float64)
    _, ev32, evl32, _ = ref_step(arch, opt, batch, torch.float32, sd=m32.state_dict(), training=False)

    # --- pin the restatement against the real reference ------------------------------------
    for dt, lg, ls, g in ((torch.float32, lg32, ls32, g32), (torch.float64, lg64, ls64, g64)):
        st = R.clone_state(init, dt)
        res = R.run_step(st, opt, batch, training=True)
        tol = 2e-5 if dt == torch.float32 else 1e-11
        assert rel(res['logits'], lg) < tol, ('logits', name, dt, rel(res['logits'], lg))
        assert abs(float(res['loss']) - float(ls)) < tol * max(1.0, abs(float(ls)))
        if dt == torch.float64:
            for k, gr in res['grads'].items():
                assert rel(gr, g[k]) < 1e-9 or float(g[k].abs().max()) < 1e-12, (k, rel(gr, g[k]))
            after = m64.state_dict()
            for k in after:
                if 'running' in k:
                    assert rel(st[k], after[k]) < 1e-12, k
        print('  restatement == reference (%s, %s): logits %.2e' % (name, dt, rel(res['logits'], lg)))

    noise = {k: float((g32[k].double() - g64[k]).norm()) for k in g32}     # absolute L2 (some true grads are 0)
    fx = dict(
        name=name, opt=dict(vars(opt)), seed=0,
        init_checksum=checksum(init),
        logits32=lg32, logits64=lg64, loss32=float(ls32), loss64=float(ls64),
        eval_logits32=ev32, eval_loss32=float(evl32),
        grad64_sample={k: sample(v).to(torch.float64) for k, v in g64.items()},
        grad64_norm={k: float(v.double().norm()) for k, v in g64.items()},
        grad64_absmax={k: float(v.double().abs().max()) for k, v in g64.items()},
        grad32_noise=noise,
        bn_after={k: sample(v) for k, v in m32.state_dict().items() if 'running' in k or 'num_batches' in k},
        bn_after64={k: sample(v) for k, v in m64.state_dict().items() if 'running' in k},
        sizes=dict(N=batch.x.shape[0], E=batch.edge.shape[0], B=int(batch.bbox_idx[-1]) + 1),
    )
    torch.save(fx, os.path.join(OUT, 'model_%s.pt' % name))
    print('  wrote model_%s.pt  loss=%.6f  noise(max)=%.2e' % (name, float(ls32), max(noise[k] / max(float(g64[k].norm()), 1e-30) for k in noise if float(g64[k].norm()) > 1e-9)))


def conv_fixture(arch, name, N, E, Cin, C, seed, with_weight=False, zero_edges=False):
    """One reference GraphConv('attr_edge_gp2') + explicit upstream grads, fp64 (and fp32 forward)."""
    import gcn_lib.sparse as gl
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    conv = gl.GraphConv(Cin, C, 'attr_edge_gp2')
    for p in conv.parameters():      # non-trivial affine / biases so that every term is exercised
        p.data = torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.5)
    sd = {k: v.clone() for k, v in conv.state_dict().items()}
    x = torch.randn(N, Cin, generator=g)
    xn = torch.randn(N, Cin, generator=g)
    if zero_edges:
        E = 0
    edge = torch.randint(0, N, (E, 2), generator=g)
    if E > 8:
        edge[1] = edge[0]                     # duplicate edge
        edge[2, 1] = edge[2, 0]               # self loop
        edge[:, 1][edge[:, 1] == N - 1] = 0   # node N-1 has no incoming edge
    attr = torch.randn(E, 4, generator=g)
    w = torch.rand(E, generator=g) if with_weight else None
    go = torch.randn(N, C, generator=g)
    gn = torch.randn(N, C, generator=g)
    res = {}
    for dt in (torch.float64, torch.float32):
        c = gl.GraphConv(Cin, C, 'attr_edge_gp2')
        c.load_state_dict(sd)
        c = c.to(dt).train()
        xx = x.to(dt).requires_grad_(True)
        xxn = xn.to(dt).requires_grad_(True)
        out, on = c(xx, edge.t(), None if w is None else w.to(dt), attr.to(dt), x_node=xxn)
        if E == 0 and dt == torch.float32:
            pass
        (out * go.to(dt)).sum().add((on * gn.to(dt)).sum()).backward()
        tag = '64' if dt == torch.float64 else '32'
        res['out' + tag], res['xnode' + tag] = out.detach(), on.detach()
        c_after = {k: v.clone() for k, v in c.state_dict().items() if 'running' in k}
        c.eval()
        eo, en = c(xx.detach(), edge.t(), None if w is None else w.to(dt), attr.to(dt), x_node=xxn.detach())
        if dt == torch.float64:      # fp64 is the anchor; fp32 only keeps the forward outputs
            res['dx64'], res['dxnode64'] = xx.grad, xxn.grad
            res['dparams64'] = {k: p.grad for k, p in c.named_parameters()}
            res['buffers_after64'] = c_after
            res['eval_out64'], res['eval_xnode64'] = eo.detach(), en.detach()
        # restatement check
        st = {('c.' + k): v.clone().to(dt) if v.is_floating_point() else v.clone() for k, v in sd.items()}
        st = {k.replace('c.gconv', 'c'): v for k, v in st.items()}
        ro, rn = R.gp2_conv(st, 'c', xx.detach(), xxn.detach(), edge.t(), attr.to(dt), True,
                            None if w is None else w.to(dt))
        tol = 1e-11 if dt == torch.float64 else 2e-5
        assert rel(ro, res['out' + tag]) < tol and rel(rn, res['xnode' + tag]) < tol, name
    fx = dict(name=name, N=N, E=E, Cin=Cin, C=C, state=sd, x=x, x_node=xn, edge=edge, attr=attr,
              edge_weight=w, grad_out=go, grad_xnode=gn, **res)
    torch.save(fx, os.path.join(OUT, 'gp2conv_%s.pt' % name))
    print('  wrote gp2conv_%s.pt' % name)


def scatter_fixture():
    from torch_scatter import scatter
    g = torch.Generator().manual_seed(7)
    cases = []
    for (M, Fd, S, sorted_idx) in ((40, 8, 6, True), (64, 33, 9, False), (5, 3, 8, True), (1, 4, 1, True)):
        src = torch.randn(M, Fd, generator=g, dtype=torch.float64)
        idx = torch.randint(0, S, (M,), generator=g)
        if sorted_idx:
            idx = idx.sort().values
        if M > 8:
            src[3] = src[2]                       # an exact tie between two rows
            idx[3] = idx[2]
        go = torch.randn(int(idx.max()) + 1, Fd, generator=g, dtype=torch.float64)
        case = dict(src=src, index=idx, grad=go)
        for red in ('mean', 'max'):
            s = src.clone().requires_grad_(True)
            o = scatter(s, idx, dim=0, reduce=red)
            (o * go).sum().backward()
            case[red] = o.detach()
            case['d' + red] = s.grad
            # restatement
            fn = R.scatter_mean if red == 'mean' else R.scatter_max
            s2 = src.clone().requires_grad_(True)
            o2 = fn(s2, idx, int(idx.max()) + 1)
            (o2 * go).sum().backward()
            assert torch.equal(o2, o) and torch.equal(s2.grad, s.grad)
        cases.append(case)
    torch.save(cases, os.path.join(OUT, 'scatter.pt'))
    print('  wrote scatter.pt')


def main():
    os.makedirs(OUT, exist_ok=True)
    arch = ref_loader.load()
    print('reference loaded from', ref_loader.REF_ROOT)
    # configs[0]: toy 128-node graph, 1 GraphConv block, 3 classes
    model_fixture(arch, 'toy', synth.toy_batch(), synth.make_opt(**synth.CONFIGS['toy'][1]))
    # configs[1] shape at reduced size (2 graphs x 400 nodes / 1600 edges)
    model_fixture(arch, 'floorplans_small', synth.floorplans_batch(graphs=2, n=400, e=1600, seed=1),
                  synth.make_opt(n_classes=17))
    # configs[2] shape at reduced size (ragged proposals)
    model_fixture(arch, 'diagrams_small', synth.diagrams_batch(graphs=2, n=300, e=900, seed=2),
                  synth.make_opt(n_classes=22))
    # a 3-block / n_blocks_out=2 variant (feature-map selection, architecture...py:60)
    model_fixture(arch, 'floorplans_3blk', synth.floorplans_batch(graphs=1, n=320, e=1200, seed=3),
                  synth.make_opt(n_classes=17, n_blocks=3, n_blocks_out=2))
    conv_fixture(arch, 'head', N=64, E=256, Cin=5, C=64, seed=11)
    conv_fixture(arch, 'block', N=64, E=256, Cin=64, C=64, seed=12)
    conv_fixture(arch, 'block_weighted', N=50, E=170, Cin=64, C=64, seed=13, with_weight=True)
    conv_fixture(arch, 'block_sparse', N=70, E=23, Cin=64, C=64, seed=14)
    scatter_fixture()
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print('golden total %.1f KB' % (tot / 1024))


if __name__ == '__main__':
    main()
