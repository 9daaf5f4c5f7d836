"""Debug aid: per-tensor errors of GraphConv('attr_edge_gp2') forward + backward against the reference goldens
(never stops at the first mismatch).  python tools/bwd_debug.py [fixture names...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
from util import load_golden, max_rel, l2_rel  # noqa: E402


def run(name):
    from yolat_vectorgraphicsrecognition_b200.gcn_lib.sparse import GraphConv
    fx = load_golden('gp2conv_%s.pt' % name)
    conv = GraphConv(fx['Cin'], fx['C'], 'attr_edge_gp2')
    conv.load_state_dict(fx['state'])
    conv = conv.cuda().train()
    x = fx['x'].detach().cuda().requires_grad_(True)
    xn = fx['x_node'].detach().cuda().requires_grad_(True)
    w = fx['edge_weight'].cuda() if fx['edge_weight'] is not None else None
    out, on = conv(x, fx['edge'].t().cuda(), w, fx['attr'].cuda(), x_node=xn)
    print('[%s] N=%d E=%d  fwd out %.2e  xnode %.2e' % (name, x.shape[0], fx['edge'].shape[0], max_rel(out, fx['out64']),
                                                        max_rel(on, fx['xnode64'])))
    (out * fx['grad_out'].cuda()).sum().add((on * fx['grad_xnode'].cuda()).sum()).backward()
    torch.cuda.synchronize()
    print('   dx %.2e  dxnode %.2e' % (l2_rel(x.grad, fx['dx64']), l2_rel(xn.grad, fx['dxnode64'])))
    for k, p in conv.named_parameters():
        ref = fx['dparams64'][k]
        if float(ref.abs().max()) < 1e-9:
            print('   %-28s zero-grad bias: max|g| %.2e' % (k, float(p.grad.abs().max())))
        else:
            print('   %-28s l2rel %.2e   (|ref| %.3e)' % (k, l2_rel(p.grad, ref), float(ref.norm())))


if __name__ == '__main__':
    for n in (sys.argv[1:] or ['head', 'block', 'block_weighted', 'block_sparse']):
        run(n)
