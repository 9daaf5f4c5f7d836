"""2-rank NCCL check of the data-parallel gradient exchange (needs 2 GPUs, skipped otherwise):
`dp.OverlappedGradSync` -- gradients written into one flat buffer by the backward kernels, the classifier-head bucket
all-reduced under the rest of the backward, the remainder after it -- must leave in every p.grad the MEAN of the two
ranks' single-rank gradients, eagerly and when both collectives are captured in the step's CUDA graph."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(150, exit=True)     # a stalled collective dumps every thread's stack and exits
    from yolat_vectorgraphicsrecognition_b200 import synth, dp
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        opt = synth.make_opt(n_classes=17)
        torch.manual_seed(0)
        model = arch.SparseCADGCN(opt).to(dev).train()
        crit = arch.DetectionLoss(opt)
        shards = [synth.floorplans_batch(graphs=1, n=640, e=2560, seed=1000 + r).to(dev) for r in range(world)]
        # reference: both shards on this rank without any exchange (fresh BN buffers each time: deepcopy of the model)
        import copy
        per = []
        for r in range(world):
            m = copy.deepcopy(model)
            crit(m(shards[r], None), shards[r])['loss'].backward()
            per.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]))
        want = sum(per) / world
        sync = dp.OverlappedGradSync(model)
        m2 = model
        for p in m2.parameters():
            p.grad = None
        crit(m2(shards[rank], None), shards[rank])['loss'].backward()
        flat = sync.finish().clone()
        assert not sync.copy_mode
        err_eager = float((flat - want).abs().max() / want.abs().max())
        overlapped, total = sync.overlapped_bytes, 4 * sync.numel
        sync.close()
        # the same exchange captured in the step graph: one bucket reduced at the end of the backward (NCCL work issued
        # asynchronously from autograd hooks INSIDE a capture hung on this stack; hooks + capture need async_op=False or
        # no hooks at all -- tools/nccl_capture_probe.py), still without a gather copy: the views are the gradients
        sync = dp.OverlappedGradSync(model, overlap=False)
        step = GraphedStep(model, crit, extra=sync.finish)
        step(shards[rank])
        torch.cuda.synchronize()
        err_graph = float((sync.flat - want).abs().max() / want.abs().max())
        adopted = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(sync.params, sync.views))
        sync.close()
        step.release()                # captured NCCL work keeps the communicator referenced: drop the graphs first
        del step
        # ... and the overlapped schedule captured: the hook forks a side stream for the head bucket (a parallel branch
        # of the step graph), finish() joins it
        sync = dp.OverlappedGradSync(model, side_stream=True)
        step = GraphedStep(model, crit, extra=sync.finish)
        step(shards[rank])
        step(shards[rank])
        torch.cuda.synchronize()
        err_side = float((sync.flat - want).abs().max() / want.abs().max())
        adopted = adopted and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(sync.params, sync.views))
        sync.close()
        step.release()
        del step
        # ... and the symmetric-memory exchange (one in-place multimem / two-shot kernel over the flat buffer), eager and
        # captured -- when the box cannot set it up every rank falls back to NCCL and the numbers must still be right
        sync = dp.OverlappedGradSync(model, symmetric=True, side_stream=True)      # head bucket on a side stream + remainder
        symm = bool(sync.symmetric)
        for p in model.parameters():
            p.grad = None
        crit(model(shards[rank], None), shards[rank])['loss'].backward()
        err_symm_eager = float((sync.finish() - want).abs().max() / want.abs().max())
        step = GraphedStep(model, crit, extra=sync.finish)
        step(shards[rank])
        step(shards[rank])
        torch.cuda.synchronize()
        err_symm = float((sync.flat - want).abs().max() / want.abs().max())
        symm_op, symm_error = sync._symm_op, sync._symm_error
        sync.close()
        step.release()
        del step
        sync = dp.OverlappedGradSync(model, symmetric=True, overlap=False)             # one call after the backward
        step = GraphedStep(model, crit, extra=sync.finish)
        step(shards[rank])
        torch.cuda.synchronize()
        err_symm_single = float((sync.flat - want).abs().max() / want.abs().max())
        if rank == 0:
            torch.save(dict(err_eager=err_eager, err_graph=err_graph, err_side=err_side, adopted=adopted,
                            overlapped=overlapped, total=total, symm=symm, symm_op=symm_op, symm_error=symm_error,
                            err_symm=max(err_symm, err_symm_single), err_symm_eager=err_symm_eager), out)
        sync.close()
        step.release()
        del step
    except BaseException:
        import traceback
        with open('%s.err%d' % (out, rank), 'w') as f:
            f.write(traceback.format_exc())
        raise
    finally:
        torch.cuda.synchronize()
        faulthandler.cancel_dump_traceback_later()
        faulthandler.dump_traceback_later(30, exit=True)
        dist.destroy_process_group()
        faulthandler.cancel_dump_traceback_later()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_overlapped_grad_sync_nccl_world2(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / 'res.pt')
    try:
        mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    except Exception:
        import glob
        for f in glob.glob(out + '.err*'):
            print(f, open(f).read())
        if not os.path.exists(out):
            raise
    res = torch.load(out)
    assert res['adopted']
    assert res['err_eager'] < 1e-6 and res['err_graph'] < 1e-6 and res['err_side'] < 1e-6, res
    assert res['overlapped'] > 0.8 * res['total']
    print('symmetric memory:', res['symm'], res['symm_op'], res['symm_error'])
    assert res['err_symm'] < 1e-6 and res['err_symm_eager'] < 1e-6, res
