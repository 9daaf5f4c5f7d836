"""Fused Adam for the hot path's parameters.

Drop-in for `torch.optim.Adam(model.parameters(), lr=opt.lr, weight_decay=opt.weight_decay)` followed by
`optimizer.step()` every iteration (cad_recognition/train.py:212, :284): same update rule (L2 penalty folded into
the gradient, bias-corrected moments, eps added after the square root), same `param_groups` surface for
`torch.optim.lr_scheduler.StepLR` (train.py:214) and the same `state_dict()` layout (`state[i] = {step, exp_avg,
exp_avg_sq}`), so `load_pretrained_optimizer` (utils/ckpt_util.py) keeps working.  The update itself is two kernel
launches over a chunk table (csrc/adam.cu) instead of ~100, reads its step counter from the device and is therefore
capturable: pass `FusedAdam.step` as `GraphedStep(extra=...)` to make it part of the replayed step.

No CPU path: parameters must live on a CUDA device.
"""
import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError('invalid Adam hyper-parameters')
        super(FusedAdam, self).__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._groups = None        # per param group: flat moment buffers, device step state, cached chunk tables

    # ------------------------------------------------------------------------------------------------
    def _prepare(self):
        lib = _lib.lib()
        chunk = lib.yolat_adam_chunk()
        self._groups = []
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.requires_grad]
            if not ps:
                self._groups.append(None)
                continue
            dev = ps[0].device
            if dev.type != 'cuda':
                raise _lib.YolatError('FusedAdam needs CUDA parameters (no CPU path)')
            for p in ps:
                if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                    raise _lib.YolatError('FusedAdam: parameters must be contiguous fp32 tensors on one device')
            n = sum(p.numel() for p in ps)
            m = torch.zeros(n, dtype=torch.float32, device=dev)
            v = torch.zeros(n, dtype=torch.float32, device=dev)
            state = torch.zeros(3, dtype=torch.float64, device=dev)
            off, views = 0, []
            for p in ps:
                views.append((m[off:off + p.numel()].view_as(p), v[off:off + p.numel()].view_as(p)))
                off += p.numel()
            # the torch-visible state: tensors aliasing the flat buffers, `step` aliasing the device counter
            for p, (mv, vv) in zip(ps, views):
                self.state[p] = {'step': state[0:1].view(()), 'exp_avg': mv, 'exp_avg_sq': vv}
            # chunk geometry is fixed by the parameter shapes; only the gradient addresses change between gradient sets
            # (the two input sets of a double-buffered GraphedStep own two sets of gradient tensors).  Tables are
            # written into pre-allocated pinned buffers and copied with an async H2D copy, which is legal inside a
            # CUDA-graph capture (the gradients of a captured backward only exist once the capture is running).
            counts, base = [], []
            off = 0
            for i, p in enumerate(ps):
                for c in range(0, p.numel(), chunk):
                    counts.append(min(chunk, p.numel() - c))
                    base.append((i, c, off + c))
                off += p.numel()
            count = torch.tensor(counts, dtype=torch.int32).to(dev)
            pool = [(torch.empty(len(counts), 4, dtype=torch.int64).pin_memory(),
                     torch.empty(len(counts), 4, dtype=torch.int64, device=dev)) for _ in range(self.MAX_GRAD_SETS)]
            self._groups.append(dict(params=ps, m=m, v=v, state=state, count=count, base=base, pool=pool, tables={}))

    MAX_GRAD_SETS = 4

    def _table(self, g):
        """Device chunk table for the current gradient tensors, cached per set of gradient addresses."""
        key = tuple(p.grad.data_ptr() for p in g['params'])
        hit = g['tables'].get(key)
        if hit is not None:
            return hit
        if len(g['tables']) >= len(g['pool']):
            g['tables'].clear()                       # gradient tensors were re-allocated (eager training): start over
        pinned, dev = g['pool'][len(g['tables'])]
        for p in g['params']:
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                raise _lib.YolatError('FusedAdam: gradients must be contiguous fp32 tensors')
        ps = g['params']
        mp, vp = g['m'].data_ptr(), g['v'].data_ptr()
        rows = [(ps[i].data_ptr() + 4 * c, ps[i].grad.data_ptr() + 4 * c, mp + 4 * o, vp + 4 * o) for i, c, o in g['base']]
        pinned.copy_(torch.tensor(rows, dtype=torch.int64))       # (addresses are below 2^63: bit-exact in int64)
        dev.copy_(pinned, non_blocking=True)
        g['tables'][key] = dev
        return dev

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._groups is None:
            self._prepare()
        lib = _lib.lib()
        for group, g in zip(self.param_groups, self._groups):
            if g is None:
                continue
            if any(p.grad is None for p in g['params']):
                raise _lib.YolatError('FusedAdam.step: every parameter of the group needs a gradient')
            table, count, n = self._table(g), g['count'], len(g['base'])
            b1, b2 = group['betas']
            _lib.check(lib.yolat_adam_step(table.data_ptr(), count.data_ptr(), n, g['state'].data_ptr(), float(group['lr']),
                                       float(b1), float(b2), float(group['eps']), float(group['weight_decay']),
                                       float(grad_scale), torch.cuda.current_stream(g['state'].device).cuda_stream), 'yolat_adam_step')
        return loss

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam / FusedAdam state dict: moments are copied into the flat buffers, the step count
        into the device counter."""
        if self._groups is None:
            self._prepare()
        groups = state_dict['param_groups']
        for group, saved in zip(self.param_groups, groups):
            for k in ('lr', 'betas', 'eps', 'weight_decay'):
                if k in saved:
                    group[k] = saved[k]
        idx = 0
        for g, group in zip(self._groups, self.param_groups):
            for p in group['params']:
                st = state_dict['state'].get(idx)
                idx += 1
                if st is None or g is None or p not in self.state:
                    continue
                self.state[p]['exp_avg'].copy_(st['exp_avg'])
                self.state[p]['exp_avg_sq'].copy_(st['exp_avg_sq'])
                g['state'][0] = float(st['step'])
