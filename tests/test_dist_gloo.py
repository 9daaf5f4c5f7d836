"""CPU, world_size 2, gloo: the data-parallel gradient exchange of SURVEY.md section 8(e).  Each rank runs
the oracle step on its own shard of the batch and all-reduces the flat gradient buffer; the result must
equal the mean of the per-shard gradients computed in one process (BN statistics stay rank-local)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restatement as R
from yolat_vectorgraphicsrecognition_b200 import synth, dp
from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_grads(rank):
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt)
    st = R.clone_state(model.state_dict(), torch.float32)
    batch = synth.floorplans_batch(graphs=1, n=160, e=640, seed=1000 + rank)
    res = R.run_step(st, opt, batch, training=True)
    return model, res['grads']


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model, grads = _shard_grads(rank)
        for k, p in model.named_parameters():
            p.grad = grads[k].clone() if grads[k] is not None else None
        fg = dp.FlatGradients(model.parameters())
        flat = fg.all_reduce_mean()
        assert flat.numel() == 1613329
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(fg.params, fg.views))
        if rank == 0:
            torch.save(flat.clone(), out)
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2(tmp_path):
    out = str(tmp_path / 'flat.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    flat = torch.load(out)
    per = []
    threads = torch.get_num_threads()
    torch.set_num_threads(1)          # same summation order as the single-threaded workers
    for r in range(2):
        model, grads = _shard_grads(r)
        per.append(torch.cat([(grads[k] if grads[k] is not None else torch.zeros_like(p)).reshape(-1)
                              for k, p in model.named_parameters()]))
    torch.set_num_threads(threads)
    want = (per[0] + per[1]) / 2
    assert torch.allclose(flat, want, rtol=0, atol=1e-7 * float(want.abs().max()))


def _worker_overlap(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model, grads = _shard_grads(rank)
        sync = dp.OverlappedGradSync(model, symmetric='auto')     # gloo / CPU: the symmetric-memory path must decline
        assert not sync.symmetric and sync._symm_op is None
        # layout: model.parameters() order, the classifier head (final first in backward) is the contiguous tail
        n_head = sum(p.numel() for k, p in model.named_parameters() if k.startswith('prediction_cls'))
        assert sync.numel == 1613329 and sync.numel - sync.split == n_head and sync.overlapped_bytes == 4 * n_head
        # two steps; on the CPU the gradients come from the oracle, so autograd cannot adopt the flat views: the object
        # must notice (hooks do not fire, pointers differ) and fall back to copy-then-reduce with the same result
        for step in range(2):
            for k, p in model.named_parameters():
                p.grad = grads[k].clone() * (step + 1) if grads[k] is not None else None
            flat = sync.finish()
            assert sync.copy_mode
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(sync.params, sync.views))
        if rank == 0:
            torch.save(flat.clone(), out)
        sync.close()
    finally:
        dist.destroy_process_group()


def test_overlapped_grad_sync_world2(tmp_path):
    out = str(tmp_path / 'flat2.pt')
    mp.spawn(_worker_overlap, args=(2, _free_port(), out), nprocs=2, join=True)
    flat = torch.load(out)
    per = []
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    for r in range(2):
        model, grads = _shard_grads(r)
        per.append(torch.cat([(grads[k] if grads[k] is not None else torch.zeros_like(p)).reshape(-1)
                              for k, p in model.named_parameters()]))
    torch.set_num_threads(threads)
    want = (per[0] + per[1])           # second step: gradients were doubled, mean of two ranks
    assert torch.allclose(flat, want, rtol=0, atol=2e-7 * float(want.abs().max()))
