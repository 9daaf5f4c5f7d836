"""2-GPU probe: under which settings can the overlapped gradient all-reduce (NCCL work issued from autograd hooks) be
captured into the step graph?  python tools/nccl_capture_probe.py <capture_error_mode> <TORCH_NCCL_ASYNC_ERROR_HANDLING>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mode, aeh = sys.argv[1], sys.argv[2]
variant = sys.argv[3] if len(sys.argv) > 3 else 'overlap'
os.environ['TORCH_NCCL_ASYNC_ERROR_HANDLING'] = aeh
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def worker(rank, world, port):
    from yolat_vectorgraphicsrecognition_b200 import synth, dp
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).to(dev).train()
    crit = arch.DetectionLoss(opt)
    b = synth.floorplans_batch(graphs=1, n=640, e=2560, seed=1000 + rank).to(dev)
    sync = dp.OverlappedGradSync(model, overlap=variant != 'single', async_early=variant == 'overlap')
    try:
        step = GraphedStep(model, crit, extra=sync.finish, capture_error_mode=mode)
        for _ in range(3):
            loss = step(b)
        torch.cuda.synchronize()
        if rank == 0:
            print('RESULT mode=%s aeh=%s variant=%s: OK loss %.5f copy_mode=%s' % (mode, aeh, variant, float(loss), sync.copy_mode), flush=True)
    except Exception as e:
        if rank == 0:
            import traceback
            print('RESULT mode=%s aeh=%s variant=%s: FAILED %s' % (mode, aeh, variant, str(e).splitlines()[0][:200]), flush=True)
            traceback.print_exc()
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


if __name__ == '__main__':
    import socket
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(2, port), nprocs=2, join=True)
