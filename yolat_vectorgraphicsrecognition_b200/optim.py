"""Fused Adam for the hot path's parameters.

Drop-in for `torch.optim.Adam(model.parameters(), lr=opt.lr, weight_decay=opt.weight_decay)` followed by
`optimizer.step()` every iteration (cad_recognition/train.py:212, :284): same update rule (L2 penalty folded into
the gradient, bias-corrected moments, eps added after the square root), same `param_groups` surface for
`torch.optim.lr_scheduler.StepLR` (train.py:214) and the same `state_dict()` layout (`state[i] = {step, exp_avg,
exp_avg_sq}`), so `load_pretrained_optimizer` (utils/ckpt_util.py) keeps working.  The update itself is two kernel
launches over a chunk table (csrc/adam.cu) instead of ~100, reads its step counter from the device and is therefore
capturable: pass `FusedAdam.step` as `GraphedStep(extra=...)` to make it part of the replayed step.  Learning rate,
betas, eps and weight decay live in device memory (`sync_hyperparams`), so a captured step follows a scheduler.

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
            state = torch.zeros(9, dtype=torch.float64, device=dev)   # step | derived | hyper-parameters
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
            # Chunk tables: one (pinned, device) pair per set of gradient addresses.  A pair consumed under CUDA-graph
            # capture is PERMANENT (the graph re-executes its pinned -> device copy on every replay, so neither buffer
            # may ever be rewritten); eager steps cycle through a small ring whose pinned side is only rewritten after
            # the event recorded behind its previous copy has completed.
            def pair():
                return (torch.empty(len(counts), 4, dtype=torch.int64).pin_memory(),
                        torch.empty(len(counts), 4, dtype=torch.int64, device=dev))
            captured = [pair() for _ in range(self.MAX_CAPTURED_SETS)]
            eager = [pair() + (torch.cuda.Event(),) for _ in range(self.EAGER_RING)]
            hyper_host = torch.zeros(6, dtype=torch.float64).pin_memory()
            self._groups.append(dict(params=ps, m=m, v=v, state=state, count=count, base=base, captured=captured,
                                     n_captured=0, eager=eager, eager_keys=[None] * self.EAGER_RING, eager_next=0,
                                     tables={}, hyper_host=hyper_host, hyper_event=None, hyper_last=None))

    MAX_CAPTURED_SETS = 32     # gradient-address sets that CUDA graphs may hold (2 input slots x 16 batch shapes)
    EAGER_RING = 2

    def prepare(self):
        """Allocate the flat moment buffers, the device state and the (pinned) table pool now -- call before capturing
        `step` in a CUDA graph (GraphedStep does): page-locked allocations are not legal while a stream captures."""
        if self._groups is None:
            self._prepare()

    def _rows(self, g):
        for p in g['params']:
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                raise _lib.YolatError('FusedAdam: gradients must be contiguous fp32 tensors')
        ps = g['params']
        mp, vp = g['m'].data_ptr(), g['v'].data_ptr()
        rows = [(ps[i].data_ptr() + 4 * c, ps[i].grad.data_ptr() + 4 * c, mp + 4 * o, vp + 4 * o) for i, c, o in g['base']]
        return torch.tensor(rows, dtype=torch.int64)       # (addresses are below 2^63: bit-exact in int64)

    def _table(self, g):
        """Device chunk table for the current gradient tensors, cached per set of gradient addresses."""
        key = tuple(p.grad.data_ptr() for p in g['params'])
        hit = g['tables'].get(key)
        if hit is not None:
            return hit
        dev = g['state'].device
        if torch.cuda.is_current_stream_capturing():
            if g['n_captured'] >= len(g['captured']):
                raise _lib.YolatError('FusedAdam: more than %d gradient sets captured in CUDA graphs' % len(g['captured']))
            pinned, table = g['captured'][g['n_captured']]
            g['n_captured'] += 1
            pinned.copy_(self._rows(g))
            table.copy_(pinned, non_blocking=True)        # a memcpy node: re-executed (same bytes) on every replay
            g['tables'][key] = table                      # never evicted
            return table
        i = g['eager_next']
        g['eager_next'] = (i + 1) % len(g['eager'])
        pinned, table, ev = g['eager'][i]
        if g['eager_keys'][i] is not None:
            g['tables'].pop(g['eager_keys'][i], None)
            ev.synchronize()                              # the previous copy out of this pinned buffer has executed
        pinned.copy_(self._rows(g))
        table.copy_(pinned, non_blocking=True)
        ev.record(torch.cuda.current_stream(dev))
        g['eager_keys'][i] = key
        g['tables'][key] = table
        return table

    def sync_hyperparams(self, grad_scale=1.0):
        """Write lr / betas / eps / weight_decay / grad_scale of every param group into its device state (one 48-byte
        asynchronous copy, only when a value changed).  `step()` calls it when it runs eagerly; a step captured in a
        CUDA graph reads the device values at replay time, so call this (GraphedStep does) before each replay."""
        if self._groups is None:
            self._prepare()
        for group, g in zip(self.param_groups, self._groups):
            if g is None:
                continue
            b1, b2 = group['betas']
            vals = (float(group['lr']), float(b1), float(b2), float(group['eps']), float(group['weight_decay']),
                    float(grad_scale))
            if vals == g['hyper_last']:
                continue
            dev = g['state'].device
            if g['hyper_event'] is not None:
                g['hyper_event'].synchronize()            # the previous copy has read the pinned buffer
            g['hyper_host'].copy_(torch.tensor(vals, dtype=torch.float64))
            g['state'][3:9].copy_(g['hyper_host'], non_blocking=True)
            if g['hyper_event'] is None:
                g['hyper_event'] = torch.cuda.Event()
            g['hyper_event'].record(torch.cuda.current_stream(dev))
            g['hyper_last'] = vals

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._groups is None:
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise _lib.YolatError('FusedAdam.step captured before prepare(): pass optimizer= to GraphedStep or call '
                                      'optimizer.prepare() and optimizer.sync_hyperparams() before the capture')
            self._prepare()
        lib = _lib.lib()
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyperparams(grad_scale)
        elif any(g is not None and g['hyper_last'] is None for g in self._groups):
            raise _lib.YolatError('FusedAdam.step captured before sync_hyperparams(): call prepare() + sync_hyperparams() '
                                  'first (GraphedStep does)')
        for group, g in zip(self.param_groups, self._groups):
            if g is None:
                continue
            if any(p.grad is None for p in g['params']):
                raise _lib.YolatError('FusedAdam.step: every parameter of the group needs a gradient')
            table, count, n = self._table(g), g['count'], len(g['base'])
            _lib.check(lib.yolat_adam_step_dev(table.data_ptr(), count.data_ptr(), n, g['state'].data_ptr(),
                                               torch.cuda.current_stream(g['state'].device).cuda_stream), 'yolat_adam_step_dev')
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
