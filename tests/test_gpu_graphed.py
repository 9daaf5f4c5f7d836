"""CUDA-graph replay of the training step must be indistinguishable from the eager step: same loss, same
gradients (the kernels are deterministic, so bit-identical), same BatchNorm running-buffer trajectory."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup():
    from yolat_vectorgraphicsrecognition_b200 import synth
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).cuda().train()
    return synth, arch, opt, model


def test_graphed_step_equals_eager_step():
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    synth, arch, opt, model = _setup()
    ref_model = copy.deepcopy(model)
    crit = arch.DetectionLoss(opt)
    batches = [synth.floorplans_batch(graphs=2, n=400, e=1600, seed=s) for s in (1, 2, 3)]

    step = GraphedStep(model, crit)
    for b in batches:                       # same shape signature: one capture, then replays
        loss_g = step(b.pin_memory())
        for p in ref_model.parameters():
            p.grad = None
        out = ref_model(b, None)
        loss_e = crit(out, b)['loss']
        loss_e.backward()
        torch.cuda.synchronize()
        assert float(loss_g.detach()) == float(loss_e.detach())
        assert torch.equal(step.last_logits, out[0])
        for (k, p), q in zip(model.named_parameters(), ref_model.parameters()):
            assert torch.equal(p.grad, q.grad), k
        for (k, a), bb in zip(model.state_dict().items(), ref_model.state_dict().values()):
            assert torch.equal(a, bb), k
    assert len(step._graphs) == 1


def test_graphed_step_recaptures_on_new_shape_and_trains():
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    synth, arch, opt, model = _setup()
    crit = arch.DetectionLoss(opt)
    optim = torch.optim.Adam(model.parameters(), lr=1e-3)
    step = GraphedStep(model, crit)
    b1 = synth.floorplans_batch(graphs=1, n=320, e=1200, seed=5)
    b2 = synth.floorplans_batch(graphs=2, n=320, e=1200, seed=6)
    losses = []
    for it in range(6):
        loss = step(b1 if it % 2 == 0 else b2)
        optim.step()
        losses.append(float(loss.detach()))
    assert len(step._graphs) == 2
    assert losses[4] < losses[0] and losses[5] < losses[1]      # the static .grad tensors feed the optimizer


def test_packed_batch_prefetch_matches_plain_replay():
    """PackedBatch (one pinned buffer, one H2D copy) + GraphedStep.prefetch (double-buffered inputs on a copy stream)
    give the same loss / gradients as the plain per-field path, step after step, with changing data."""
    from yolat_vectorgraphicsrecognition_b200.batch import PackedBatch
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    synth, arch, opt, model = _setup()
    ref_model = arch.SparseCADGCN(opt).cuda().train()
    ref_model.load_state_dict(model.state_dict())
    crit = arch.DetectionLoss(opt)
    step, ref_step = GraphedStep(model, crit), GraphedStep(ref_model, crit)
    batches = [synth.floorplans_batch(graphs=1, n=640, e=2560, seed=s) for s in (1, 2, 3, 4)]
    packed = [PackedBatch.from_batch(b) for b in batches]
    assert packed[0].host.is_pinned()
    step.prefetch(packed[0])
    for i, (pb, b) in enumerate(zip(packed, batches)):
        loss = step(pb)
        if i + 1 < len(packed):
            step.prefetch(packed[i + 1])
        ref = ref_step(b.to('cuda'))
        assert abs(float(loss) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
        g = model.prediction_cls[2][0].weight.grad
        gr = ref_model.prediction_cls[2][0].weight.grad
        assert float((g - gr).abs().max()) <= 1e-6 * float(gr.abs().max()) + 1e-12


def test_captured_graphs_survive_workspace_growth():
    """A graph captured for a small batch shape keeps replaying correctly after the shared scratch workspace has grown
    (larger shapes, eval / predict passes) and after other tensors have been allocated: its workspace block is retired,
    never freed (ADVICE r1 / VERDICT r1 weak #7).  Small / large / small / ... with bit-identical gradients on every
    step against an eager model."""
    from yolat_vectorgraphicsrecognition_b200 import _lib
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    synth, arch, opt, model = _setup()
    ref_model = copy.deepcopy(model)
    crit = arch.DetectionLoss(opt)
    small = synth.floorplans_batch(graphs=1, n=320, e=1200, seed=5)
    large = synth.floorplans_batch(graphs=4, n=2000, e=8000, seed=6)
    huge = synth.floorplans_batch(graphs=8, n=4000, e=16000, seed=7)
    step = GraphedStep(model, crit)
    dev = next(model.parameters()).device
    junk = []
    for it, b in enumerate([small, large, small, huge, small, large, small]):
        if it == 3:                       # an eager no-grad pass on an even larger shape grows the workspace again
            with torch.no_grad():
                model(huge, None)
                ref_model(huge, None)
        loss_g = step(b)
        junk.append(torch.full((1 << 22,), float('nan'), device=dev))     # lands on any block the allocator got back
        for p in ref_model.parameters():
            p.grad = None
        out = ref_model(b, None)
        loss_e = crit(out, b)['loss']
        loss_e.backward()
        torch.cuda.synchronize()
        assert float(loss_g.detach()) == float(loss_e.detach()), it
        for (k, p), q in zip(model.named_parameters(), ref_model.parameters()):
            assert torch.equal(p.grad, q.grad), (it, k)
    assert len(step._graphs) == 3
    assert _lib.workspace.retired_bytes(dev) > 0          # the smaller blocks are still owned


def test_gradients_are_written_into_the_flat_buffer():
    """dp.OverlappedGradSync: the backward kernels write every parameter gradient straight into views of one flat
    buffer and autograd adopts them as p.grad (no gather copy before the all-reduce), eagerly and under graph replay;
    values equal a plain run bit for bit."""
    from yolat_vectorgraphicsrecognition_b200 import dp
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    synth, arch, opt, model = _setup()
    ref_model = copy.deepcopy(model)
    crit = arch.DetectionLoss(opt)
    b = synth.floorplans_batch(graphs=1, n=640, e=2560, seed=1).to('cuda')
    crit(ref_model(b, None), b)['loss'].backward()
    sync = dp.OverlappedGradSync(model)
    try:
        crit(model(b, None), b)['loss'].backward()
        flat = sync.finish()
        assert not sync.copy_mode
        lo, hi = flat.data_ptr(), flat.data_ptr() + 4 * flat.numel()
        for (k, p), q in zip(model.named_parameters(), ref_model.parameters()):
            assert lo <= p.grad.data_ptr() < hi, k
            assert torch.equal(p.grad, q.grad), k
        want = torch.cat([q.grad.reshape(-1) for q in ref_model.parameters()])
        assert torch.equal(flat, want)
        step = GraphedStep(model, crit, extra=sync.finish)
        for _ in range(2):
            step(b)
        assert not sync.copy_mode and torch.equal(sync.flat, want)
    finally:
        sync.close()
