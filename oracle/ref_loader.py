"""Import the UNMODIFIED reference model from /root/reference through oracle/shims (build container only).

/root/reference does not exist on the GPU box: nothing reachable from `-m gpu` tests, smoke() or
bench.py may call this module.  It is used by `oracle/make_golden.py` and by CPU-side tests that are
skipped when the reference tree is absent.
"""
import os
import sys

import torch

REF_ROOT = '/root/reference'
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'cad_recognition'))


_loaded = None


def load():
    """Returns the reference module `architecture3cc_rpn_gp_iter2` (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('/root/reference is not present on this machine')
    for p in (os.path.join(REF_ROOT, 'cad_recognition'), REF_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # the reference hard-codes .cuda() (architecture3cc_rpn_gp_iter2.py:107-115,371): identity on a CPU box
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    import architecture3cc_rpn_gp_iter2 as arch
    _loaded = arch
    return arch


def unload():
    """Drop the reference modules and shims from sys.modules / sys.path (so product tests see a clean state)."""
    global _loaded
    _loaded = None
    for name in list(sys.modules):
        root = name.split('.')[0]
        if root in ('architecture3cc_rpn_gp_iter2', 'gcn_lib', 'utils', 'torch_scatter', 'torch_geometric',
                    'torch_cluster', 'thop', 'fvcore', 'h5py'):
            mod = sys.modules[name]
            f = getattr(mod, '__file__', '') or ''
            if f.startswith(REF_ROOT) or f.startswith(_SHIMS):
                del sys.modules[name]
    for p in (os.path.join(REF_ROOT, 'cad_recognition'), REF_ROOT, _SHIMS):
        while p in sys.path:
            sys.path.remove(p)


def build_reference_model(opt, seed=0, dtype=torch.float32):
    arch = load()
    torch.manual_seed(seed)
    model = arch.SparseCADGCN(opt)
    crit = arch.DetectionLoss(opt)
    return model.to(dtype), crit
