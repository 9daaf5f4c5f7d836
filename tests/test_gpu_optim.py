"""Fused multi-tensor Adam (csrc/adam.cu, optim.FusedAdam) against torch.optim.Adam on the same parameters and
gradients: same update rule (L2 weight decay folded into the gradient), same state-dict layout, capturable."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 14), (64,), (1024, 128), (17, 256), (3,), (2049,)]
    return [torch.randn(*s, generator=g).cuda().requires_grad_() for s in shapes]


@pytest.mark.parametrize('wd', [0.0, 1e-4])
def test_fused_adam_matches_torch_adam(wd):
    from yolat_vectorgraphicsrecognition_b200.optim import FusedAdam
    a, b = _params(0), _params(0)
    ref = torch.optim.Adam(a, lr=1e-3, weight_decay=wd)
    opt = FusedAdam(b, lr=1e-3, weight_decay=wd)
    g = torch.Generator().manual_seed(1)
    for it in range(5):
        for pa, pb in zip(a, b):
            gr = torch.randn(pa.shape, generator=g).cuda()
            pa.grad, pb.grad = gr.clone(), gr.clone()
        ref.step()
        opt.step()
    for pa, pb in zip(a, b):
        assert float((pa.detach() - pb.detach()).abs().max()) <= 1e-5 * max(1.0, float(pa.detach().abs().max()))
    sa, sb = ref.state_dict(), opt.state_dict()
    assert sorted(sa['state'].keys()) == sorted(sb['state'].keys())
    for k in sa['state']:
        assert float(sb['state'][k]['step']) == 5.0
        ea, eb = sa['state'][k]['exp_avg'], sb['state'][k]['exp_avg']
        assert float((ea - eb).abs().max()) <= 1e-5 * max(1e-3, float(ea.abs().max()))
        va, vb = sa['state'][k]['exp_avg_sq'], sb['state'][k]['exp_avg_sq']
        assert float((va - vb).abs().max()) <= 1e-5 * max(1e-3, float(va.abs().max()))


def test_fused_adam_resumes_from_a_torch_adam_checkpoint_and_follows_the_scheduler():
    from yolat_vectorgraphicsrecognition_b200.optim import FusedAdam
    a, b = _params(2), _params(2)
    ref = torch.optim.Adam(a, lr=2e-3, weight_decay=1e-5)
    g = torch.Generator().manual_seed(3)
    grads = [[torch.randn(p.shape, generator=g).cuda() for p in a] for _ in range(6)]
    for it in range(3):
        for p, gr in zip(a, grads[it]):
            p.grad = gr.clone()
        ref.step()
    with torch.no_grad():
        for pa, pb in zip(a, b):
            pb.copy_(pa)
    opt = FusedAdam(b, lr=1.0)                      # hyper-parameters come from the checkpoint
    opt.load_state_dict(ref.state_dict())
    sched_a = torch.optim.lr_scheduler.StepLR(ref, 2, 0.5)
    sched_b = torch.optim.lr_scheduler.StepLR(opt, 2, 0.5)
    for it in range(3, 6):
        for pa, pb, gr in zip(a, b, grads[it]):
            pa.grad, pb.grad = gr.clone(), gr.clone()
        ref.step(); opt.step()
        sched_a.step(); sched_b.step()
    assert opt.param_groups[0]['lr'] == ref.param_groups[0]['lr']
    for pa, pb in zip(a, b):
        assert float((pa.detach() - pb.detach()).abs().max()) <= 1e-5 * max(1.0, float(pa.detach().abs().max()))


def test_fused_adam_is_capturable_in_the_step_graph():
    """GraphedStep(optimizer=FusedAdam): forward + loss + backward + Adam replayed as one graph.  The capture's warm-up
    steps must not update anything (step counter == replays == BatchNorm batches tracked), and the captured step must
    follow a learning-rate change made after the capture (hyper-parameters live in device memory).

    The reference is an eager torch.optim.Adam fed with the SAME gradients (copied out of the replayed step), so the
    comparison isolates the optimizer arithmetic.  (Two independently trained copies diverge chaotically at the 1e-5
    level within 8 steps: biases feeding a BatchNorm have mathematically-zero gradients, Adam turns their rounding noise
    into lr-sized steps, and those feed back through the features.)"""
    import copy
    from yolat_vectorgraphicsrecognition_b200 import synth
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    from yolat_vectorgraphicsrecognition_b200.optim import FusedAdam
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).cuda().train()
    ref_model = copy.deepcopy(model)
    crit = arch.DetectionLoss(opt)
    optim = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    ref_optim = torch.optim.Adam(ref_model.parameters(), lr=1e-3, weight_decay=1e-5)
    batch = synth.floorplans_batch(graphs=1, n=640, e=2560, seed=1).to('cuda')

    step = GraphedStep(model, crit, optimizer=optim)
    losses = []
    for it in range(8):
        if it == 4:                                   # what StepLR does between epochs (train.py:214)
            optim.param_groups[0]['lr'] = 2.5e-4
            ref_optim.param_groups[0]['lr'] = 2.5e-4
        losses.append(float(step(batch).detach()))
        for pr, pm in zip(ref_model.parameters(), model.parameters()):
            pr.grad = pm.grad.detach().clone()        # the gradients the replay just consumed
        ref_optim.step()
    assert losses[-1] < losses[0], losses
    assert float(optim.state_dict()['state'][0]['step']) == 8.0      # no hidden updates during the capture warm-up
    for k, buf in model.named_buffers():
        if k.endswith('num_batches_tracked'):
            assert int(buf) == 8, (k, int(buf))                      # ... and no hidden BatchNorm updates either
    for (k, a), b in zip(model.named_parameters(), ref_model.parameters()):
        if a.dim() < 2:
            continue      # zero-gradient biases: m / sqrt(v) of pure rounding noise, decided by the last bit of either optimizer
        assert float((a.detach() - b.detach()).abs().max()) <= 1e-6 * max(1.0, float(b.detach().abs().max())), k
