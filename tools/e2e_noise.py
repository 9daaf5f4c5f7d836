"""Diagnostic: end-to-end gradient error of the full config-2 step against the fp64 restatement, worst tensors first.
Run with YOLAT_GEMM=simt (exact fp32 SIMT GEMMs) to separate the 3xTF32 products' contribution to ReLU / arg-max flips."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import synth  # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch  # noqa: E402
from oracle import restatement as R  # noqa: E402

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
opt = synth.make_opt(n_classes=17)
torch.manual_seed(0)
model = arch.SparseCADGCN(opt)
st = R.clone_state(model.state_dict(), torch.float64)
batch = synth.floorplans_batch(graphs=graphs, seed=1)
ref = R.run_step(st, opt, batch, training=True)
ref32 = R.run_step(R.clone_state(model.state_dict(), torch.float32), opt, batch, training=True)
model = model.cuda().train()
out = model(batch, None)
loss = arch.DetectionLoss(opt)(out, batch)['loss']
loss.backward()
rows = []
for k, p in model.named_parameters():
    g = ref['grads'][k]
    if float(g.abs().max()) < 1e-12:
        continue
    rows.append((float((p.grad.double().cpu() - g).norm() / g.norm()), float((ref32['grads'][k].double() - g).norm() / g.norm()), k))
rows.sort(reverse=True)
print('mode GEMM=%s EDGE_BWD=%s  logits err %.2e' % (os.environ.get('YOLAT_GEMM', 'tc'), os.environ.get('YOLAT_EDGE_BWD', 'recompute'),
                                                     float((out[0].double().cpu() - ref['logits']).abs().max() / ref['logits'].abs().max())))
for r in rows[:8]:
    print('   %-44s ours %.2e   ref32 %.2e' % (r[2], r[0], r[1]))
print('   median ours %.2e' % rows[len(rows) // 2][0])
