"""CUDA-graph replay of the training step (forward + DetectionLoss + backward).

The config-2 step is ~150 short kernels; launched eagerly from Python it is bound by the host (~2.1 ms of
enqueue for ~1.6 ms of device work).  Every C-ABI entry point is stream-ordered, allocation-free and
sync-free, so the whole step can be captured once per batch *shape* and replayed with one launch:

    step = GraphedStep(model, DetectionLoss(opt))
    loss = step(batch)        # batch: CPU (pinned) or CUDA tensors, like train.py:270-283
    optimizer.step()          # p.grad tensors are static across replays

With a `batch.PackedBatch` (one pinned buffer) the inputs reach the device as ONE copy, and
`step.prefetch(next_batch)` stages the next batch on a copy stream into a second input set while the current
step runs (what a prefetching data loader does), so the host->device traffic leaves the critical path.

A batch with a new shape signature (N, E, B) is captured on first use (3 eager warm-up steps, then the
capture); graphs are kept per signature, so the padded / bucketed batches of a data loader re-use them.
BatchNorm running statistics and `num_batches_tracked` are updated by the kernels themselves, so replay
keeps the reference's training semantics (cad_recognition/train.py:263-286).
"""
from types import SimpleNamespace

import torch

from .batch import PackedBatch

_FIELDS = ('x', 'bbox_idx', 'edge', 'bbox', 'e_attr', 'labels')


class GraphedStep(object):

    def __init__(self, model, criterion, warmup=3, extra=None, optimizer=None, capture_error_mode='global'):
        """`extra(loss)` (optional) is called inside the captured region after backward -- e.g. a fused
        optimizer step or the data-parallel gradient all-reduce -- and must be capture-safe.  It is NOT run during the
        eager warm-up steps of a capture (they would be hidden optimizer updates / extra collectives); if its owner
        (`extra.__self__`) has `prepare()` it is called once here (allocations that are illegal under capture), and
        `sync_hyperparams()` before every capture and replay (device-resident learning rate of optim.FusedAdam).
        `optimizer=` names that owner explicitly (and makes `extra` its step when no `extra` is given).
        With a collective inside `extra`, every rank must meet a new batch shape at the same step: capture is a
        collective event.  `capture_error_mode` is passed to torch.cuda.graph ('thread_local' when autograd hooks issue
        NCCL work from the autograd thread during capture: dp.OverlappedGradSync)."""
        if optimizer is not None and extra is None:       # GraphedStep(model, crit, optimizer=FusedAdam(...))
            extra = lambda loss: optimizer.step()
        self.model, self.criterion, self.warmup, self.extra = model, criterion, int(warmup), extra
        self.capture_error_mode = capture_error_mode
        self._extra_owner = optimizer if optimizer is not None else getattr(extra, '__self__', None)
        if hasattr(self._extra_owner, 'prepare'):
            self._extra_owner.prepare()
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.device = next(model.parameters()).device
        if self.device.type != 'cuda':
            raise RuntimeError('GraphedStep needs the model on a CUDA device (no CPU path)')
        self._graphs = {}
        self._copy_stream = None
        self._staged = {}          # id(batch) -> entry whose inputs are in flight on the copy stream
        self._next_slot = 0

    @staticmethod
    def signature(batch):
        if isinstance(batch, PackedBatch):
            return batch.signature() + (bool(getattr(batch, 'offsets_pending', False)),)
        return tuple((f, tuple(getattr(batch, f).shape), getattr(batch, f).dtype) for f in _FIELDS)

    def _sync_extra(self):
        if hasattr(self._extra_owner, 'sync_hyperparams'):
            self._extra_owner.sync_hyperparams()

    def _eager(self, static, run_extra=True):
        if getattr(static, '_apply_offsets', False):     # batch.collate(...).defer_offsets(): train.py:238-258 on the device
            from .batch import apply_offsets
            apply_offsets(static)
        out = self.model(static, None)
        loss = self.criterion(out, static)['loss']
        loss.backward()
        if run_extra and self.extra is not None:
            self.extra(loss)
        return loss, out[0]

    @staticmethod
    def _stage(entry, batch):
        """Enqueue the host->device copies of `batch` into the entry's static inputs on the current stream:
        one copy for a PackedBatch, one per field otherwise."""
        if entry.buf is not None:
            entry.buf.copy_(batch.host, non_blocking=True)
        else:
            for f in _FIELDS:
                getattr(entry.static, f).copy_(getattr(batch, f), non_blocking=True)

    def _capture(self, batch):
        if isinstance(batch, PackedBatch):
            buf, static = batch.device_twin(self.device)
            static._apply_offsets = bool(getattr(batch, 'offsets_pending', False))
        else:
            buf = None
            static = SimpleNamespace(**{f: torch.empty_like(getattr(batch, f), device=self.device) for f in _FIELDS})
        entry = SimpleNamespace(buf=buf, static=static, ready=None, done=torch.cuda.Event())
        self._stage(entry, batch)
        # warm up on a side stream: lazy initialisation (cudaFuncSetAttribute, workspace growth, cuda context
        # pieces) must happen outside the capture; gradients are created here and stay the same tensors
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        restage = bool(getattr(static, '_apply_offsets', False))      # the device offsets update the inputs in place
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                for p in self.params:
                    p.grad = None
                if restage:
                    self._stage(entry, batch)
                self._eager(static, run_extra=False)
            if restage:
                self._stage(entry, batch)
            if hasattr(self._extra_owner, 'drain'):      # e.g. dp.OverlappedGradSync: collectives issued by the warm-up hooks
                self._extra_owner.drain()
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._sync_extra()
        # NOTE: the warm-up steps are real training-mode forwards + backwards (without `extra`): they update the BN
        # running buffers the same way `warmup` extra iterations on this batch would; _entry() restores them, so a
        # capture leaves parameters, optimizer state and buffers untouched.
        graph = torch.cuda.CUDAGraph()
        for p in self.params:
            p.grad = None
        with torch.cuda.graph(graph, capture_error_mode=self.capture_error_mode):
            loss, logits = self._eager(static)
        entry.graph, entry.loss, entry.logits = graph, loss, logits
        entry.grads = [p.grad for p in self.params]
        return entry

    def _entry(self, batch, slot):
        """The captured graph (and its static inputs) for this batch shape; `slot` selects one of two input sets
        so that the next batch can be staged while the current one is still being read."""
        key = (self.signature(batch), slot)
        entry = self._graphs.get(key)
        if entry is None:
            state = {k: v.clone() for k, v in self.model.state_dict().items() if 'running_' in k or 'num_batches' in k}
            grads = [p.grad for p in self.params]      # a capture (e.g. from prefetch) must not disturb the caller's grads
            entry = self._capture(batch)
            for p, g in zip(self.params, grads):
                p.grad = g
            self._graphs[key] = entry
            with torch.no_grad():     # undo the BN-buffer updates of the warm-up steps (capture itself runs nothing)
                sd = self.model.state_dict()
                for k, v in state.items():
                    sd[k].copy_(v)
            entry.fresh = True
        return entry

    def release(self):
        """Drop every captured graph (and its static inputs).  Call it before `dist.destroy_process_group()` when `extra`
        issued collectives: a captured NCCL kernel keeps the communicator referenced until its graph is destroyed."""
        torch.cuda.synchronize(self.device)
        self._staged.clear()
        for entry in self._graphs.values():
            entry.graph = None
        self._graphs.clear()
        import gc
        gc.collect()
        torch.cuda.synchronize(self.device)

    def prefetch(self, batch):
        """Stage `batch` (host tensors, ideally a pinned PackedBatch) for the NEXT call on a copy stream, into the
        input set the step in flight is not using: its host->device copy overlaps the current step's kernels."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        slot = self._next_slot
        self._next_slot ^= 1
        entry = self._entry(batch, slot)
        cur = torch.cuda.current_stream(self.device)
        if getattr(entry, 'fresh', False):
            entry.fresh = False                   # the capture just staged this very batch on the current stream
            entry.ready = torch.cuda.Event()
            entry.ready.record(cur)
        else:
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(entry.done)      # the last replay that read this input set has finished
                self._stage(entry, batch)
                entry.ready = torch.cuda.Event()
                entry.ready.record(self._copy_stream)
        self._staged[id(batch)] = entry
        return entry

    def __call__(self, batch):
        cur = torch.cuda.current_stream(self.device)
        entry = self._staged.pop(id(batch), None)
        if entry is not None:
            cur.wait_event(entry.ready)
        else:
            entry = self._entry(batch, 0)
            if getattr(entry, 'fresh', False):
                entry.fresh = False
            else:
                self._stage(entry, batch)
        for p, g in zip(self.params, entry.grads):
            p.grad = g
        self._sync_extra()
        entry.graph.replay()
        entry.done.record(cur)
        self.last_logits = entry.logits
        return entry.loss
