"""ctypes binding of libyolat_b200.so -- the C ABI declared in include/yolat_b200.h.

There is deliberately NO fallback: if the shared library is missing or a call returns a negative
status, this raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C yolat_vectorgraphicsrecognition_b200/csrc`.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('YOLAT_B200_LIB') or os.path.join(_HERE, 'libyolat_b200.so')   # override: kernel experiments

i64, i32, f32p, vp = C.c_int64, C.c_int, C.c_void_p, C.c_void_p


class YolatBn(C.Structure):
    _fields_ = [('w', vp), ('b', vp), ('running_mean', vp), ('running_var', vp), ('num_batches_tracked', vp)]


class Gp2Params(C.Structure):
    _fields_ = [('w1', vp), ('b1', vp), ('bn1', YolatBn),
                ('w2', vp), ('b2', vp), ('bn2', YolatBn),
                ('wr', vp), ('br', vp),
                ('wn', vp), ('bnode', vp), ('bnn', YolatBn)]


class Gp2Grads(C.Structure):
    _fields_ = [(n, vp) for n in ('w1', 'b1', 'bn1_w', 'bn1_b', 'w2', 'b2', 'bn2_w', 'bn2_b', 'wr', 'br',
                                  'wn', 'bnode', 'bnn_w', 'bnn_b')]


# name -> (restype, argtypes); mirrors include/yolat_b200.h one to one
_SIGS = {
    'yolat_abi_version': (C.c_int, []),
    'yolat_status_string': (C.c_char_p, [C.c_int]),
    'yolat_last_cuda_error': (C.c_char_p, []),
    'yolat_launch_count': (i64, []),
    'yolat_prof_enable': (C.c_int, [C.c_int]),
    'yolat_prof_read': (C.c_int, [C.c_int, C.POINTER(i64), C.POINTER(C.c_double)]),
    'yolat_graph_ints': (i64, [i64, i64]),
    'yolat_graph_build': (C.c_int, [vp, i64, i64, i64, i64, vp, vp]),
    'yolat_graph_error_ptr': (vp, [vp, i64, i64]),
    'yolat_graph_rowptr': (vp, [vp, i64, i64]),
    'yolat_graph_src': (vp, [vp, i64, i64]),
    'yolat_graph_eid': (vp, [vp, i64, i64]),
    'yolat_segments_ints': (i64, [i64, i64]),
    'yolat_segments_build': (C.c_int, [vp, i64, i64, vp, vp]),
    'yolat_gp2_tape_floats': (i64, [i64, i64, i32, i32, i32]),
    'yolat_gp2_tape_floats_mode': (i64, [i64, i64, i32, i32, i32, i32]),
    'yolat_gp2_fwd_ws_floats': (i64, [i64, i64, i32, i32, i32]),
    'yolat_gp2_bwd_ws_floats': (i64, [i64, i64, i32, i32, i32]),
    'yolat_gp2_fwd': (C.c_int, [C.POINTER(Gp2Params), i32, i32, i32, vp, i64, vp, i64, vp, vp, vp, i64, i64, i32,
                                vp, i64, vp, i64, vp, i64, vp, i64, vp]),
    'yolat_gp2_bwd': (C.c_int, [C.POINTER(Gp2Params), C.POINTER(Gp2Grads), i32, i32, i32, vp, i64, vp, i64, vp, vp, vp,
                                i64, i64, i32, vp, i64, vp, i64, vp, i64, vp, i64, i32, vp, vp, i64, vp]),
    'yolat_edge1_tape_floats': (i64, [i64, i64, i32, i32]),
    'yolat_edge1_ws_floats': (i64, [i64, i64, i32, i32]),
    'yolat_edge1_fwd': (C.c_int, [vp, vp, C.POINTER(YolatBn), i32, i32, vp, i64, vp, vp, vp, i64, i64, i32, vp, i64, vp, i64,
                                  vp, i64, vp]),
    'yolat_edge1_bwd': (C.c_int, [vp, C.POINTER(YolatBn), i32, i32, vp, i64, vp, vp, vp, i64, i64, i32, vp, i64, vp, i64, i32,
                                  vp, vp, vp, vp, vp, vp, i64, vp]),
    'yolat_gemm_ws_floats': (i64, [i32, i64, i64, i64]),
    'yolat_gemm': (C.c_int, [i32, vp, i64, vp, i64, vp, i64, i64, i64, i64, vp, i32, vp, i64, vp]),
    'yolat_mlp_tape_floats': (i64, [i64, i32, i32, i32]),
    'yolat_mlp_ws_floats': (i64, [i64, i32, i32, i32]),
    'yolat_mlp_fwd': (C.c_int, [vp, i64, i64, i32, vp, vp, i32, C.POINTER(YolatBn), i32, vp, i64, vp, i64, vp, i64, vp]),
    'yolat_mlp_bwd': (C.c_int, [vp, i64, i64, i32, vp, i32, C.POINTER(YolatBn), i32, vp, i64, vp, i64, i32, vp, vp, vp,
                                vp, vp, vp, i64, vp]),
    'yolat_segment_mean_fwd': (C.c_int, [vp, i64, i64, i32, vp, i64, vp, i64, vp]),
    'yolat_segment_mean_bwd': (C.c_int, [vp, i64, i64, i32, vp, i64, vp, i64, i32, vp]),
    'yolat_segment_max_fwd': (C.c_int, [vp, i64, i64, i32, vp, i64, vp, i64, vp, vp]),
    'yolat_segment_max_bwd': (C.c_int, [vp, i64, i64, i32, i64, vp, vp, i64, i32, vp]),
    'yolat_fusemax_tape_floats': (i64, [i64, i32, i32, i64]),
    'yolat_fusemax_ws_floats': (i64, [i64, i32, i32, i64]),
    'yolat_fusemax_fwd': (C.c_int, [vp, i64, i64, i32, vp, vp, i32, C.POINTER(YolatBn), i32, vp, i64, vp, i64, vp, i64,
                                    vp, i64, vp]),
    'yolat_fusemax_bwd': (C.c_int, [vp, i64, i64, i32, vp, i32, C.POINTER(YolatBn), i32, vp, i64, vp, i64, vp, i64, i32,
                                    vp, vp, vp, vp, vp, vp, i64, vp]),
    'yolat_softmax_xent_fwd': (C.c_int, [vp, i64, i64, i32, vp, vp, vp, vp, i64, vp]),
    'yolat_softmax_xent_bwd': (C.c_int, [vp, i64, i32, vp, vp, vp, i64, vp]),
    'yolat_expand_ranges': (C.c_int, [vp, vp, i64, i64, vp, vp]),
    'yolat_slice_graph_ints': (i64, [i64, i64]),
    'yolat_slice_graph': (C.c_int, [vp, i64, vp, i64, vp, i64, i64, vp, vp, vp, vp, vp]),
    'yolat_batch_offsets': (C.c_int, [vp, i64, vp, i64, vp, i64, vp]),
    'yolat_proposals_ws_bytes': (i64, [vp]),
    'yolat_proposals_count': (C.c_int, [vp, vp, i64, vp, vp]),
    'yolat_proposals_fill': (C.c_int, [vp, vp, i64, vp, vp]),
    'yolat_adam_chunk': (C.c_int, []),
    'yolat_adam_step': (C.c_int, [vp, vp, i64, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.c_double, vp]),
    'yolat_adam_step_dev': (C.c_int, [vp, vp, i64, vp, vp]),
}

_lib = None


class YolatError(RuntimeError):
    pass


def lib():
    """The loaded CDLL; raises (never falls back) if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise YolatError('%s is missing: build the sm_100a extension first '
                             '(python -c "import __graft_entry__ as g; g.build()")' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)        # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(status, what=''):
    if status != 0:
        l = lib()
        msg = l.yolat_status_string(status).decode()
        cuda = l.yolat_last_cuda_error().decode() if status == -3 else ''
        raise YolatError('yolat_b200 %s failed: %s (%d) %s' % (what, msg, status, cuda))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise YolatError('yolat_b200 has no CPU path: tensors must live on a CUDA device (got %s)' % t.device)


def f32c(t):
    """fp32 + contiguous view of a tensor (no copy when already so)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def f32rows(t):
    """fp32 matrix whose rows are contiguous (row pitch >= row length): column slices of a wider matrix -- what the
    backward of `torch.cat(..., dim=1)` hands to the producers of its inputs -- pass through without a copy; the C-ABI
    takes the pitch as `ld`."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return t
    return t if t.is_contiguous() else t.contiguous()


class _Workspace(object):
    """One grow-only fp32 scratch buffer per device, shared by all calls on the same stream order.

    A CUDA graph captured by `GraphedStep` bakes the buffer's device address into its kernel nodes, so a buffer that
    has been handed out is NEVER released: when a larger request arrives the old block is retired (kept referenced for
    the life of the process) and a bigger one becomes current.  Graphs captured earlier keep replaying into their own,
    still-owned block; nothing else can be allocated on top of it."""

    def __init__(self):
        self.buf = {}
        self.retired = {}

    def get(self, n_floats, device):
        n_floats = max(int(n_floats), 64)
        b = self.buf.get(device)
        if b is None or b.numel() < n_floats:
            if b is not None:
                self.retired.setdefault(device, []).append(b)
            b = torch.empty(int(n_floats * 1.25), dtype=torch.float32, device=device)
            self.buf[device] = b
        return b

    def retired_bytes(self, device):
        return sum(t.numel() * 4 for t in self.retired.get(device, []))


workspace = _Workspace()
