"""Data-parallel plumbing for the hot path.

Graph batches shard across ranks (one process per GPU, no data-path collective); the only exchange is ONE
all-reduce of the flat gradient buffer per step (1 613 329 fp32 = 6.45 MB at the README config).  BatchNorm
statistics stay rank-local (DDP-default semantics).  The reference has no working multi-GPU path
(`DataParallel(SparseDeepGCN(...))` is a NameError, cad_recognition/train.py:204-205), so this is new.
Backend-agnostic: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_graphs(n_graphs, world_size, rank, edges_per_graph=None):
    """Contiguous split of the batch's graphs across ranks, balanced by edge count when given.
    Returns the [begin, end) graph range of `rank`."""
    if edges_per_graph is None:
        base, rem = divmod(n_graphs, world_size)
        begin = rank * base + min(rank, rem)
        return begin, begin + base + (1 if rank < rem else 0)
    total = float(sum(edges_per_graph))
    bounds, acc, r = [0], 0.0, 1
    for g, e in enumerate(edges_per_graph):
        acc += e
        while r < world_size and acc >= total * r / world_size - 1e-9:
            bounds.append(g + 1)
            r += 1
    while len(bounds) < world_size + 1:
        bounds.append(n_graphs)
    bounds[-1] = n_graphs
    return bounds[rank], bounds[rank + 1]


class FlatGradients(object):
    """One contiguous fp32 buffer holding every parameter's gradient, all-reduced in a single call."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=p0.dtype, device=p0.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def gather(self):
        """Copy p.grad into the flat buffer (missing grads count as zero)."""
        src = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, src)
        return self.flat

    def all_reduce_mean(self, group=None):
        """flat <- mean over ranks; p.grad re-pointed at the flat views.  Returns the flat buffer."""
        self.gather()
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(1.0 / world)
        for p, v in zip(self.params, self.views):
            p.grad = v
        return self.flat


class OverlappedGradSync(object):
    """Gradient averaging that overlaps the backward pass (last row; section 8e).

    * Every parameter's gradient lives in ONE flat buffer from the start: the backward kernels write into views of it
      (ops.set_grad_arena), autograd adopts those views as `p.grad` -- no gather copy before the collective, and a
      fused optimizer sees one set of gradient addresses for every captured graph.
    * The buffer is laid out in `model.parameters()` order, so the parameters whose gradients are final FIRST in the
      backward pass -- the classifier head `prediction_cls.*`, 82 % of the 6.45 MB at the README config -- form its
      contiguous tail.  A post-accumulate hook counts them and launches their all-reduce (asynchronously, on the
      process group's own stream) the moment the last one is written, while the fusion block / GraphConv backward
      (~85 % of the backward's runtime) is still running.
    * `finish()` -- call it after `loss.backward()`, or pass it as `GraphedStep(extra=...)` so that both collectives
      are captured into the step graph -- reduces the remaining head of the buffer, joins both collectives and scales
      by 1 / world.  Only that second, small all-reduce is exposed.

    `side_stream=True` is the capturable form of the same overlap: the hook forks a side stream off the backward's
    stream, issues the early all-reduce there synchronously (no Work handle outlives the hook) and `finish()` joins the
    side stream back -- inside a CUDA-graph capture this becomes a parallel branch of the step graph.  (Asynchronous NCCL
    work issued from autograd hooks inside a capture fails with cudaErrorStreamCaptureIsolation or hangs on torch 2.11 /
    NCCL 2.28; tools/nccl_capture_probe.py and tools/nccl_hang_probe.py list what captures cleanly.)
    `overlap=False` drops the hooks: one collective over the whole buffer in `finish()`.
    `symmetric=True` (NCCL groups on one NVSwitch node) allocates the flat buffer as torch symmetric memory and reduces
    it in `finish()` with ONE in-place kernel over peer / multicast addresses (`symm_mem.multimem_all_reduce_` when the
    fabric supports multicast -- the reduction happens in the switch --, else `two_shot_all_reduce_`): a 6 MB buffer
    is latency-bound, and this path has a third of NCCL's latency and, unlike an overlapped NCCL kernel, never holds
    SMs while the persistent compute kernels (one CTA per SM) are running.  `symmetric='auto'` selects it whenever it can be
    set up; with `overlap=True` the head bucket is reduced by the same kernel on a side stream under the backward.  All ranks agree on availability with one
    all-reduce; when any rank cannot set it up, every rank silently keeps the NCCL schedule (`self.symmetric` tells).

    BatchNorm statistics stay rank-local (DDP-default semantics).  Gradients must be re-created every step
    (`zero_grad(set_to_none=True)`, what GraphedStep does): an existing `p.grad` makes autograd accumulate in place
    instead of adopting the view.  If autograd ever declines a view (verified on every eager step), the object falls
    back to the copy-then-reduce schedule of `FlatGradients`.
    """

    def __init__(self, model, group=None, early_prefixes=('prediction_cls',), overlap=True, async_early=True,
                 side_stream=False, symmetric=False):
        from . import ops
        self.overlap, self.async_early, self.side_stream = bool(overlap), bool(async_early), bool(side_stream)
        self._side, self._side_pending = None, False
        self.symmetric, self._symm_group, self._symm_op, self._symm_error = False, None, None, None
        self._avg = None
        named = [(k, p) for k, p in model.named_parameters() if p.requires_grad]
        self.params = [p for _, p in named]
        self.group = group
        p0 = self.params[0]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None
        # 'auto': measured on B200 / NVSwitch with this 6.45 MB buffer -- exposed time per step, NCCL overlapped vs
        # symmetric memory (one call after the backward) vs symmetric memory with the head bucket on a side stream:
        # 2 ranks 56 / 74 / 37 us, 8 ranks 83 / 44 / see DESIGN.md 3.5.  The overlapped symmetric schedule wins everywhere.
        if symmetric == 'auto':
            symmetric = dist.is_initialized() and self.world() >= 2
        if symmetric:
            self._try_symmetric(p0.device)
        if self.flat is None:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += p.numel()
        # the early bucket must be a contiguous tail of the parameter order
        early = [any(k.startswith(pre) for pre in early_prefixes) for k, _ in named]
        first_early = early.index(True) if True in early else len(early)
        if not all(early[first_early:]):
            first_early = len(early)              # not a tail: a single bucket at the end
        self.split = self.offsets[first_early] if first_early < len(early) else self.numel
        self.early_params = self.params[first_early:]
        self.late_params = self.params[:first_early]
        self.views = [self.flat.narrow(0, o, p.numel()).view(p.shape) for o, p in zip(self.offsets, self.params)]
        ops.set_grad_arena({p.data_ptr(): (self.flat, o, p.numel()) for o, p in zip(self.offsets, self.params)})
        self.copy_mode = False
        self.enabled = True           # False: the hooks do nothing (steps that must not communicate, e.g. an A/B timing)
        self._count = 0
        self._work = None
        if self.symmetric:
            # the two kernel calls split the buffer at a 4096-element boundary at or after the bucket boundary (the
            # multimem kernels want aligned, 16-byte-multiple slices); the few early-bucket elements below it simply
            # travel with the second call
            self._symm_split = min((self.split + 4095) // 4096 * 4096, self._symm_full.numel())
            if self._symm_split >= self._symm_full.numel() or self.split >= self.numel:
                self.overlap = False
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.early_params] if self.overlap else []
        if not self.overlap:
            self.split = self.numel           # one bucket, reduced in finish()
        self.exposed_bytes = 4 * self.split
        self.overlapped_bytes = 4 * (self.numel - self.split)

    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _try_symmetric(self, device):
        """Flat buffer in symmetric memory + rendezvous; every rank must succeed or every rank falls back."""
        if not (dist.is_initialized() and self.world() > 1 and device.type == 'cuda'
                and dist.get_backend(self.group) == 'nccl'):
            return
        ok, flat, full, op = 0, None, None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            pg = self.group if self.group is not None else dist.group.WORLD
            padded = (self.numel + 4095) // 4096 * 4096            # the kernels want 16-byte multiples per rank
            full = symm_mem.empty(padded, dtype=torch.float32, device=device)
            full.zero_()
            hdl = symm_mem.rendezvous(full, pg)
            op = 'multimem_all_reduce_' if int(getattr(hdl, 'multicast_ptr', 0) or 0) != 0 else 'two_shot_all_reduce_'
            if not hasattr(torch.ops.symm_mem, op):
                raise RuntimeError('torch.ops.symm_mem.%s is missing' % op)
            flat = full.narrow(0, 0, self.numel)
            ok = 1
        except Exception as e:        # noqa: BLE001 -- any set-up failure means "use NCCL"
            self._symm_error = '%s: %s' % (type(e).__name__, str(e).splitlines()[0][:200] if str(e) else '')
        agree = torch.tensor([ok, 1 if op == 'multimem_all_reduce_' else 0], dtype=torch.int32, device=device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
        if int(agree[0].item()) != 1:
            return
        pg = self.group if self.group is not None else dist.group.WORLD
        self._symm_full, self.flat = full, flat
        self._symm_group = pg.group_name
        self._symm_op = 'multimem_all_reduce_' if int(agree[1].item()) == 1 else 'two_shot_all_reduce_'
        self.symmetric = True

    def _op(self):
        """ncclAvg folds the 1 / world scaling into the collective (no separate kernel over the buffer); gloo has no AVG."""
        if self._avg is None:
            self._avg = dist.is_initialized() and dist.get_backend(self.group) == 'nccl'
        return dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM

    def _adopted(self, early):
        n_late = len(self.late_params)
        params, views = (self.early_params, self.views[n_late:]) if early else (self.late_params, self.views[:n_late])
        return all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, views))

    def _on_grad(self, _p):
        if not self.enabled:
            return
        self._count += 1
        if self._count < len(self.early_params) or self.copy_mode:
            return
        self._count = 0
        if not self._adopted(True):
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise RuntimeError('OverlappedGradSync: autograd did not adopt the flat-buffer views during capture')
            self.copy_mode = True
            return
        if self._work is not None:                # a step whose finish() was never called: see drain()
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise RuntimeError('OverlappedGradSync: an uncaptured collective is pending; call drain() before the capture')
            self._work.wait()
            self._work = None
        if self.world() > 1 and self.split < self.numel and self.symmetric:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.flat.device)
            self._side.wait_stream(torch.cuda.current_stream(self.flat.device))         # fork: the early grads are final
            with torch.cuda.stream(self._side):
                early = self._symm_full.narrow(0, self._symm_split, self._symm_full.numel() - self._symm_split)
                getattr(torch.ops.symm_mem, self._symm_op)(early, 'sum', self._symm_group)
                early.mul_(1.0 / self.world())
            self._side_pending = True
        elif self.world() > 1 and self.split < self.numel:
            early = self.flat.narrow(0, self.split, self.numel - self.split)
            if self.side_stream:
                if self._side is None:
                    self._side = torch.cuda.Stream(device=self.flat.device)
                self._side.wait_stream(torch.cuda.current_stream(self.flat.device))     # fork: the early grads are final
                with torch.cuda.stream(self._side):
                    dist.all_reduce(early, op=self._op(), group=self.group)
                self._side_pending = True
            else:
                self._work = dist.all_reduce(early, op=self._op(), group=self.group, async_op=self.async_early)

    def finish(self, _loss=None):
        """Join the early collective, reduce the rest, scale: p.grad = mean over ranks for every parameter."""
        self._count = 0
        if not self.enabled:              # a step that must not communicate (see `enabled`)
            return self.flat
        world = self.world()
        self._op()
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        if not self.copy_mode and not (self._adopted(False) and self._adopted(True)):
            if capturing:
                raise RuntimeError('OverlappedGradSync: autograd did not adopt the flat-buffer views during capture')
            self.copy_mode = True
        if self.copy_mode:
            if self._work is not None:
                self._work.wait()
                self._work = None
            self._join_side()
            src = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
            torch._foreach_copy_(self.views, src)
            if world > 1:
                dist.all_reduce(self.flat, op=self._op(), group=self.group)
            for p, v in zip(self.params, self.views):
                p.grad = v
        elif world > 1 and self.symmetric:
            late = self._symm_full.narrow(0, 0, self._symm_split) if self.overlap else self._symm_full
            getattr(torch.ops.symm_mem, self._symm_op)(late, 'sum', self._symm_group)
            late.mul_(1.0 / world)
            self._join_side()
            return self.flat
        elif world > 1:
            if self.split > 0:
                dist.all_reduce(self.flat.narrow(0, 0, self.split), op=self._op(), group=self.group)
            if self._work is not None:
                self._work.wait()
                self._work = None
        self._join_side()
        if world > 1 and not self._avg:
            self.flat.mul_(1.0 / world)
        return self.flat

    def _join_side(self):
        if self._side_pending:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
            self._side_pending = False

    def drain(self):
        """Join an early-bucket collective whose finish() was never called (the eager warm-up steps of a GraphedStep
        capture run the backward hooks but not `extra`).  Must run OUTSIDE a capture: a captured stream may not depend on
        uncaptured work of another stream (cudaErrorStreamCaptureIsolation)."""
        self._count = 0
        if self._work is not None:
            self._work.wait()
            self._work = None
        self._join_side()

    def close(self):
        from . import ops
        for h in self._hooks:
            h.remove()
        ops.set_grad_arena(None)
