"""CUDA-graph replay of the training step (forward + DetectionLoss + backward).

The config-2 step is ~150 short kernels; launched eagerly from Python it is bound by the host (~2.1 ms of
enqueue for ~1.6 ms of device work).  Every C-ABI entry point is stream-ordered, allocation-free and
sync-free, so the whole step can be captured once per batch *shape* and replayed with one launch:

    step = GraphedStep(model, DetectionLoss(opt))
    loss = step(batch)        # batch: CPU (pinned) or CUDA tensors, like train.py:270-283
    optimizer.step()          # p.grad tensors are static across replays

A batch with a new shape signature (N, E, B) is captured on first use (3 eager warm-up steps, then the
capture); graphs are kept per signature, so the padded / bucketed batches of a data loader re-use them.
BatchNorm running statistics and `num_batches_tracked` are updated by the kernels themselves, so replay
keeps the reference's training semantics (cad_recognition/train.py:263-286).
"""
from types import SimpleNamespace

import torch

_FIELDS = ('x', 'bbox_idx', 'edge', 'bbox', 'e_attr', 'labels')


class GraphedStep(object):

    def __init__(self, model, criterion, warmup=3, extra=None):
        """`extra(loss)` (optional) is called inside the captured region after backward -- e.g. a fused
        optimizer step or the data-parallel gradient all-reduce -- and must be capture-safe."""
        self.model, self.criterion, self.warmup, self.extra = model, criterion, int(warmup), extra
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.device = next(model.parameters()).device
        if self.device.type != 'cuda':
            raise RuntimeError('GraphedStep needs the model on a CUDA device (no CPU path)')
        self._graphs = {}

    @staticmethod
    def signature(batch):
        return tuple((f, tuple(getattr(batch, f).shape), getattr(batch, f).dtype) for f in _FIELDS)

    def _eager(self, static):
        out = self.model(static, None)
        loss = self.criterion(out, static)['loss']
        loss.backward()
        if self.extra is not None:
            self.extra(loss)
        return loss, out[0]

    def _capture(self, batch):
        static = SimpleNamespace(**{f: torch.empty_like(getattr(batch, f), device=self.device) for f in _FIELDS})
        for f in _FIELDS:
            getattr(static, f).copy_(getattr(batch, f), non_blocking=True)
        # warm up on a side stream: lazy initialisation (cudaFuncSetAttribute, workspace growth, cuda context
        # pieces) must happen outside the capture; gradients are created here and stay the same tensors
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                for p in self.params:
                    p.grad = None
                self._eager(static)
        torch.cuda.current_stream(self.device).wait_stream(side)
        # NOTE: the warm-up steps are real training-mode forwards: they update the BN running buffers the same
        # way `warmup` extra iterations on this batch would.  Restore them so capture has no side effect.
        graph = torch.cuda.CUDAGraph()
        for p in self.params:
            p.grad = None
        with torch.cuda.graph(graph):
            loss, logits = self._eager(static)
        entry = SimpleNamespace(graph=graph, static=static, loss=loss, logits=logits,
                                grads=[p.grad for p in self.params])
        return entry

    def __call__(self, batch):
        sig = self.signature(batch)
        entry = self._graphs.get(sig)
        fresh = entry is None
        if fresh:
            state = {k: v.clone() for k, v in self.model.state_dict().items() if 'running_' in k or 'num_batches' in k}
            entry = self._capture(batch)
            self._graphs[sig] = entry
            with torch.no_grad():     # undo the BN-buffer updates of the warm-up steps (capture itself runs nothing)
                sd = self.model.state_dict()
                for k, v in state.items():
                    sd[k].copy_(v)
        else:
            for f in _FIELDS:
                getattr(entry.static, f).copy_(getattr(batch, f), non_blocking=True)
        for p, g in zip(self.params, entry.grads):
            p.grad = g
        entry.graph.replay()
        self.last_logits = entry.logits
        return entry.loss
