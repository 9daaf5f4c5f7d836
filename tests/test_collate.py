"""Batch assembly: `batch.collate` against the UNMODIFIED reference `train.collate` and the per-image
offset loop of `train.train` (cad_recognition/train.py:123-171, :238-258), driven through the reference's own code.

CPU part (needs /root/reference, skipped elsewhere): the reference's train.py is imported through oracle/shims, its
`train()` loop is run for one iteration with a recording model, once on `train.collate` output and once on OUR
`batch.collate` output (a PackedBatch): slices, collated tensors and the offset tensors the model receives must be
bit-identical, and the oracle restatement of the offset loop must reproduce them from the packed slice tables.
GPU part: csrc/slicing.cu::k_batch_offsets == that restatement, and a GraphedStep on a batch with deferred (device)
offsets == the same step on host-offset tensors."""
import copy
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from util import ROOT
from oracle import batch_offsets as BO
from oracle import ref_loader


def _images(n_images=3, seed=0):
    """Per-image Data-like objects as Datasets/graph_dict3.py:966-1092 emits them (local indices)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shims'))
    from torch_geometric.data import Data
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_images):
        n_prop = int(torch.randint(2, 6, (1,), generator=g))
        sizes = torch.randint(3, 9, (n_prop,), generator=g)
        N = int(sizes.sum())
        E = int(torch.randint(N, 3 * N, (1,), generator=g))
        pos = torch.rand(N, 2, generator=g)
        out.append(Data(x=torch.cat([torch.zeros(N, 3), pos], 1), pos=pos,
                        edge=torch.randint(0, N, (E, 2), generator=g), e_attr=torch.randn(E, 4, generator=g),
                        bbox_idx=torch.repeat_interleave(torch.arange(n_prop), sizes),
                        bbox=torch.rand(n_prop, 4, generator=g), labels=torch.randint(0, 17, (n_prop,), generator=g),
                        is_super=torch.zeros(N, dtype=torch.long), stat_feats=torch.rand(n_prop, 13, generator=g),
                        roots=[SimpleNamespace(id=(i, k)) for k in range(2)], filepath='img%d.svg' % i, width=100 + i))
    return out


class _Recorder(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.seen = None

    def forward(self, data, slices):
        self.seen = {k: (data[k].clone() if torch.is_tensor(data[k]) else copy.copy(data[k])) for k in data.keys}
        self.slices = {k: v.clone() for k, v in slices.items()}
        return (self.w.sum() * 0 + 1.0,)


def _run_reference_loop(train_mod, data, slices, tmp):
    """One iteration of the reference's train() (train.py:233-321) with a recording model."""
    model = _Recorder()
    optim = torch.optim.SGD(model.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.StepLR(optim, 10)
    opt = SimpleNamespace(iter=0, arch='gcn', losses=train_mod.AverageMeter(), print_freq=10 ** 9, epoch=0, test_value=0.0,
                          best_value=0.0, writer=SimpleNamespace(add_scalar=lambda *a, **k: None), ckpt_dir=str(tmp),
                          postname='t')
    train_mod.train(model, [(data, slices)], optim, sched, lambda out, d: {'loss': out[0]}, opt)
    return model


@pytest.mark.skipif(not ref_loader.available(), reason='needs the reference tree (build container only)')
def test_collate_and_offsets_match_the_reference_loop(tmp_path):
    from yolat_vectorgraphicsrecognition_b200 import batch as B
    ref_loader.load()
    import train as ref_train          # the UNMODIFIED /root/reference/cad_recognition/train.py
    try:
        imgs = _images()
        ref_data, ref_slices = ref_train.collate(copy.deepcopy(imgs))
        pb, slices = B.collate(copy.deepcopy(imgs), pin=False)
        assert sorted(slices) == sorted(ref_slices)
        for k in ref_slices:
            assert torch.equal(slices[k], ref_slices[k]), k
        for k in ref_data.keys:
            a, b = ref_data[k], pb[k]
            assert (torch.equal(a, b) if torch.is_tensor(a) else [getattr(x, 'id', x) for x in a] ==
                    [getattr(x, 'id', x) for x in b]), k
        raw_edge, raw_bidx = pb.edge.clone(), pb.bbox_idx.clone()
        # the reference's own loop, on its own collate output and on ours (in place on the packed buffer's views)
        m_ref = _run_reference_loop(ref_train, ref_data, ref_slices, tmp_path)
        m_ours = _run_reference_loop(ref_train, pb, slices, tmp_path)
        for k in ('x', 'edge', 'e_attr', 'bbox_idx', 'bbox', 'labels'):
            assert torch.equal(m_ref.seen[k], m_ours.seen[k]), k
        assert pb.edge.data_ptr() == getattr(pb, 'edge').data_ptr() and torch.equal(pb.edge, m_ref.seen['edge'])
        # the slice tables packed with the batch reproduce the loop (oracle restatement = what the device kernel does)
        e2, b2 = BO.apply_offsets(raw_edge.numpy(), raw_bidx.numpy(), getattr(pb, B.OFFSET_FIELD).numpy())
        assert np.array_equal(e2, m_ref.seen['edge'].numpy()) and np.array_equal(b2, m_ref.seen['bbox_idx'].numpy())
    finally:
        ref_loader.unload()
        sys.modules.pop('train', None)


def test_collate_layout_without_the_reference():
    """The same contract checked against the restatement only (runs on the GPU box too): one pinned buffer, prefix-sum
    slices, per-image offsets from the packed tables."""
    from yolat_vectorgraphicsrecognition_b200 import batch as B
    imgs = _images(n_images=4, seed=3)
    pb, slices = B.collate(imgs, pin=False)
    assert pb.host.numel() == pb.nbytes and pb.edge.shape[0] == sum(i.edge.shape[0] for i in imgs)
    assert slices['edge'].tolist() == np.cumsum([0] + [i.edge.shape[0] for i in imgs]).tolist()
    e2, b2 = BO.apply_offsets(pb.edge.numpy(), pb.bbox_idx.numpy(), getattr(pb, B.OFFSET_FIELD).numpy())
    off_n = np.cumsum([0] + [i.x.shape[0] for i in imgs])
    off_b = np.cumsum([0] + [i.labels.shape[0] for i in imgs])
    want_e = np.concatenate([i.edge.numpy() + off_n[k] for k, i in enumerate(imgs)])
    want_b = np.concatenate([i.bbox_idx.numpy() + off_b[k] for k, i in enumerate(imgs)])
    assert np.array_equal(e2, want_e) and np.array_equal(b2, want_b)
    assert b2.max() + 1 == pb.bbox.shape[0] and e2.max() < pb.x.shape[0]


@pytest.mark.gpu
def test_device_offsets_match_the_restatement_and_the_host_path():
    from yolat_vectorgraphicsrecognition_b200 import batch as B, synth
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    imgs = _images(n_images=5, seed=7)
    pb, slices = B.collate(imgs)
    want_e, want_b = BO.apply_offsets(pb.edge.numpy(), pb.bbox_idx.numpy(), getattr(pb, B.OFFSET_FIELD).numpy())
    buf, ns = pb.device_twin('cuda')
    buf.copy_(pb.host)
    B.apply_offsets(ns)
    assert np.array_equal(ns.edge.cpu().numpy(), want_e) and np.array_equal(ns.bbox_idx.cpu().numpy(), want_b)
    # the whole step: device offsets (captured in the graph) == host offsets
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).cuda().train()
    ref_model = copy.deepcopy(model)
    crit = arch.DetectionLoss(opt)
    host = SimpleNamespace(x=pb.x.clone(), edge=torch.from_numpy(want_e), e_attr=pb.e_attr.clone(),
                           bbox_idx=torch.from_numpy(want_b), bbox=pb.bbox.clone(), labels=pb.labels.clone())
    loss_ref = crit(ref_model(host, None), host)['loss']
    loss_ref.backward()
    step = GraphedStep(model, crit)
    for _ in range(2):
        loss = step(pb.defer_offsets())
    torch.cuda.synchronize()
    assert float(loss.detach()) == float(loss_ref.detach())
    for (k, p), q in zip(model.named_parameters(), ref_model.parameters()):
        assert torch.equal(p.grad, q.grad), k
    # eager call on the packed batch: one upload + the same device offsets
    out = ref_model(pb, None)
    assert float(crit(out, pb)['loss'].detach()) == float(loss_ref.detach())


@pytest.mark.skipif(not ref_loader.available(), reason='needs the reference tree (build container only)')
def test_reference_train_py_binds_to_the_overlay():
    """INTEGRATION.md section 1 executed: with the alias stub in place the UNMODIFIED reference train.py imports and its
    symbols (`train.py:30` SparseCADGCN / DetectionLoss; `:201` construction; `:212` Adam over model.parameters();
    utils/ckpt_util.py strict load_state_dict) resolve to this repo's engine, whose state-dict layout is the
    reference's own.  (Running a step needs the GPU: tests/test_gpu_model.py.)"""
    import importlib
    from yolat_vectorgraphicsrecognition_b200 import synth
    names = ('gcn_lib', 'gcn_lib.sparse', 'torch_scatter', 'architecture3cc_rpn_gp_iter2')
    ref_loader.load()
    import architecture3cc_rpn_gp_iter2 as ref_arch
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    ref_sd = ref_arch.SparseCADGCN(opt).state_dict()
    saved = {n: sys.modules.get(n) for n in names + ('train',)}
    try:
        # --- the stub of INTEGRATION.md ---
        sys.modules['gcn_lib'] = importlib.import_module('yolat_vectorgraphicsrecognition_b200.gcn_lib')
        sys.modules['gcn_lib.sparse'] = importlib.import_module('yolat_vectorgraphicsrecognition_b200.gcn_lib.sparse')
        sys.modules['torch_scatter'] = importlib.import_module('yolat_vectorgraphicsrecognition_b200.torch_scatter')
        sys.modules['architecture3cc_rpn_gp_iter2'] = importlib.import_module(
            'yolat_vectorgraphicsrecognition_b200.architecture3cc_rpn_gp_iter2')
        sys.modules.pop('train', None)
        import train as ref_train
        ours = sys.modules['architecture3cc_rpn_gp_iter2']
        assert ref_train.SparseCADGCN is ours.SparseCADGCN and ref_train.DetectionLoss is ours.DetectionLoss
        torch.manual_seed(0)
        model = ref_train.SparseCADGCN(opt)                         # train.py:201
        sd = model.state_dict()
        assert list(sd) == list(ref_sd)
        for k in sd:
            assert sd[k].shape == ref_sd[k].shape and torch.equal(sd[k], ref_sd[k]), k      # same seeded init
        model.load_state_dict(ref_sd, strict=True)                 # utils/ckpt_util.py:67
        torch.optim.Adam(model.parameters(), lr=2.5e-4, weight_decay=1e-5)      # train.py:212
        assert isinstance(ref_train.DetectionLoss(opt), torch.nn.Module)
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
        ref_loader.unload()
        sys.modules.pop('train', None)
