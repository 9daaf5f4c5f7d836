"""torch.autograd.Function wrappers around the C ABI (include/yolat_b200.h).

Each Function allocates outputs and the tape with torch (PyTorch owns all memory), passes raw device
pointers + the current CUDA stream to libyolat_b200.so, and never touches the data on the host.
Reference call sites are cited next to each op.
"""
import ctypes as C

import torch

from . import _lib as L
from .graph import CSRGraph, Segments

BN, RELU, TRAINING = 1, 2, 4


def _bn_struct(weight, bias, running_mean, running_var, nbt):
    return L.YolatBn(L.ptr(weight), L.ptr(bias), L.ptr(running_mean), L.ptr(running_var), L.ptr(nbt))


# Gradient arena (dp.OverlappedGradSync): parameter storage address -> (flat buffer, offset, numel).  When a parameter
# is registered, the backward kernels write its gradient straight into that slice of ONE flat buffer (autograd's
# AccumulateGrad adopts the returned view as p.grad), so the data-parallel all-reduce needs no gather copy and a fused
# optimizer sees the same gradient addresses for every captured graph.
_GRAD_ARENA = {}


def set_grad_arena(entries):
    """entries: {param.data_ptr(): (flat, offset, numel)} or None / {} to clear."""
    _GRAD_ARENA.clear()
    if entries:
        _GRAD_ARENA.update(entries)


def _grad_for(ptr, like):
    """Gradient buffer for the parameter stored at `ptr`: a fresh view of the arena slice, or a new tensor."""
    e = _GRAD_ARENA.get(ptr)
    if e is None:
        return torch.empty_like(like, memory_format=torch.contiguous_format)
    flat, off, n = e
    return flat.narrow(0, off, n).view(like.shape)


def _empty_like_param(p):
    return _grad_for(p.data_ptr(), p)


# --------------------------------------------------------------------------------------------------
# GraphConv('attr_edge_gp2')   gcn_lib/sparse/torch_vertex.py:288-341
# --------------------------------------------------------------------------------------------------
class GP2ConvFn(torch.autograd.Function):
    """(x, x_node, attr, edge_weight, 14 parameters) -> (out, x_node_out); buffers are updated in place.
    `save_tape` = autograd is recording (torch.is_grad_enabled() at the call site): without it the forward runs the
    fully fused K-EDGE passes and writes no per-edge activation."""

    N_PARAMS = 14

    @staticmethod
    def forward(ctx, graph, training, save_tape, buffers, x, x_node, attr, edge_weight, *params):
        (w1, b1, g1, be1, w2, b2, g2, be2, wr, br, wn, bnode, gn, ben) = params
        (rm1, rv1, nbt1, rm2, rv2, nbt2, rmn, rvn, nbtn) = buffers
        L.require_cuda(x, x_node, attr, w1)
        lib = L.lib()
        x, x_node = L.f32c(x), L.f32c(x_node)
        attr = L.f32c(attr) if attr is not None else None
        ew = L.f32c(edge_weight) if edge_weight is not None else None
        N, Cin = x.shape
        Cn = x_node.shape[1]
        C_ = w2.shape[0]
        E = graph.E
        if graph.N != N:
            raise ValueError('graph was built for %d nodes, x has %d rows' % (graph.N, N))
        if E > 0 and (attr is None or tuple(attr.shape) != (E, 4)):
            raise ValueError('edge_attr must be [E, 4]')
        params = tuple(L.f32c(p) for p in params)
        (w1, b1, g1, be1, w2, b2, g2, be2, wr, br, wn, bnode, gn, ben) = params
        P = L.Gp2Params(L.ptr(w1), L.ptr(b1), _bn_struct(g1, be1, rm1, rv1, nbt1),
                        L.ptr(w2), L.ptr(b2), _bn_struct(g2, be2, rm2, rv2, nbt2),
                        L.ptr(wr), L.ptr(br),
                        L.ptr(wn), L.ptr(bnode), _bn_struct(gn, ben, rmn, rvn, nbtn))
        out = torch.empty(N, C_, dtype=torch.float32, device=x.device)
        xn_out = torch.empty(N, C_, dtype=torch.float32, device=x.device)
        # bit 1 = forward only: no per-edge activation is written when autograd will not come back for it
        mode = int(bool(training)) | (0 if (save_tape and any(ctx.needs_input_grad)) else 2)
        tape_n = lib.yolat_gp2_tape_floats_mode(N, E, Cin, Cn, C_, mode)
        ws_n = lib.yolat_gp2_fwd_ws_floats(N, E, Cin, Cn, C_)
        if ws_n < 0:
            raise L.YolatError('attr_edge_gp2: out_channels must be 32, 64 or 128 (got %d)' % C_)
        tape = torch.empty(max(tape_n, 1), dtype=torch.float32, device=x.device)
        ws = L.workspace.get(ws_n, x.device)
        L.check(lib.yolat_gp2_fwd(C.byref(P), Cin, Cn, C_, x.data_ptr(), x.stride(0), x_node.data_ptr(),
                                  x_node.stride(0), L.ptr(attr), L.ptr(ew), graph.ptr(), N, E, mode,
                                  out.data_ptr(), out.stride(0), xn_out.data_ptr(), xn_out.stride(0),
                                  tape.data_ptr(), tape.numel(), ws.data_ptr(), ws.numel(), L.stream()), 'gp2_fwd')
        ctx.graph, ctx.training, ctx.buffers = graph, int(bool(training)), buffers
        ctx.dims = (N, E, Cin, Cn, C_)
        ctx.save_for_backward(x, x_node, attr, ew, tape, *params)
        return out, xn_out

    @staticmethod
    def backward(ctx, g_out, g_xn):
        lib = L.lib()
        x, x_node, attr, ew, tape = ctx.saved_tensors[:5]
        params = ctx.saved_tensors[5:]
        (w1, b1, g1, be1, w2, b2, g2, be2, wr, br, wn, bnode, gn, ben) = params
        (rm1, rv1, nbt1, rm2, rv2, nbt2, rmn, rvn, nbtn) = ctx.buffers
        N, E, Cin, Cn, C_ = ctx.dims
        g_out, g_xn = L.f32rows(g_out), L.f32rows(g_xn)     # column slices of a cat's gradient: no copy, pitch = ld
        P = L.Gp2Params(L.ptr(w1), L.ptr(b1), _bn_struct(g1, be1, rm1, rv1, None),
                        L.ptr(w2), L.ptr(b2), _bn_struct(g2, be2, rm2, rv2, None),
                        L.ptr(wr), L.ptr(br),
                        L.ptr(wn), L.ptr(bnode), _bn_struct(gn, ben, rmn, rvn, None))
        grads = [_empty_like_param(p) for p in params]
        G = L.Gp2Grads(*[g.data_ptr() for g in grads])
        need_dx, need_dxn = ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        dx = torch.empty_like(x) if need_dx else None
        dxn = torch.empty_like(x_node) if need_dxn else None
        ws = L.workspace.get(lib.yolat_gp2_bwd_ws_floats(N, E, Cin, Cn, C_), x.device)
        L.check(lib.yolat_gp2_bwd(C.byref(P), C.byref(G), Cin, Cn, C_, x.data_ptr(), x.stride(0), x_node.data_ptr(),
                                  x_node.stride(0), L.ptr(attr), L.ptr(ew), ctx.graph.ptr(), N, E, ctx.training,
                                  g_out.data_ptr(), g_out.stride(0), g_xn.data_ptr(), g_xn.stride(0),
                                  L.ptr(dx), Cin, L.ptr(dxn), Cn, 0, tape.data_ptr(), ws.data_ptr(), ws.numel(),
                                  L.stream()), 'gp2_bwd')
        return (None, None, None, None, dx, dxn, None, None) + tuple(grads)


# --------------------------------------------------------------------------------------------------
# one-layer edge convolutions: GraphConv('edge' | 'attr_edge' | 'attr_edge_gp')   torch_vertex.py:219-286,343-484
# --------------------------------------------------------------------------------------------------
class Edge1ConvFn(torch.autograd.Function):
    """(base, x, attr, edge_weight, w1e, b1, gamma, beta) -> base + mean_{e -> i} w_e relu(bn(w1e [x_i | x_j - x_i | attr] + b1)).
    `w1e` [C, 2 Cin + 4] is the recipe's Linear weight already embedded in the GP2 column layout (the embedding is a
    differentiable torch op of the caller, so autograd maps the gradient back to the recipe's own weight)."""

    @staticmethod
    def forward(ctx, graph, training, buffers, base, x, attr, edge_weight, w1e, b1, gamma, beta):
        L.require_cuda(x, w1e, base)
        lib = L.lib()
        x, w1e, attr = L.f32c(x), L.f32c(w1e), L.f32c(attr)
        ew = L.f32c(edge_weight) if edge_weight is not None else None
        N, Cin = x.shape
        C_ = w1e.shape[0]
        E = graph.E
        if graph.N != N or tuple(w1e.shape) != (C_, 2 * Cin + 4) or tuple(attr.shape) != (E, 4):
            raise ValueError('edge conv: inconsistent shapes')
        rm, rv, nbt = buffers
        bn = _bn_struct(gamma, beta, rm, rv, nbt if training else None)
        out = L.f32c(base).clone()
        ws_n = lib.yolat_edge1_ws_floats(N, E, Cin, C_)
        if ws_n < 0:
            raise L.YolatError('edge conv: out_channels must be 32, 64 or 128 (got %d)' % C_)
        tape = torch.empty(max(lib.yolat_edge1_tape_floats(N, E, Cin, C_), 1), dtype=torch.float32, device=x.device)
        ws = L.workspace.get(ws_n, x.device)
        L.check(lib.yolat_edge1_fwd(w1e.data_ptr(), L.ptr(b1), C.byref(bn), Cin, C_, x.data_ptr(), x.stride(0), L.ptr(attr),
                                    L.ptr(ew), graph.ptr(), N, E, int(bool(training)), out.data_ptr(), out.stride(0),
                                    tape.data_ptr(), tape.numel(), ws.data_ptr(), ws.numel(), L.stream()), 'edge1_fwd')
        ctx.graph, ctx.training, ctx.buffers, ctx.has_bias = graph, int(bool(training)), buffers, b1 is not None
        ctx.dims = (N, E, Cin, C_)
        ctx.save_for_backward(x, attr, ew, w1e, gamma, beta, tape)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = L.lib()
        x, attr, ew, w1e, gamma, beta, tape = ctx.saved_tensors
        N, E, Cin, C_ = ctx.dims
        g_out = L.f32c(g_out)
        rm, rv, _ = ctx.buffers
        bn = _bn_struct(gamma, beta, rm, rv, None)
        dx = torch.empty_like(x) if ctx.needs_input_grad[4] else None
        dw = torch.empty_like(w1e)
        db = torch.empty(C_, dtype=torch.float32, device=x.device)
        dg, dbe = torch.empty_like(gamma), torch.empty_like(beta)
        ws = L.workspace.get(lib.yolat_edge1_ws_floats(N, E, Cin, C_), x.device)
        L.check(lib.yolat_edge1_bwd(w1e.data_ptr(), C.byref(bn), Cin, C_, x.data_ptr(), x.stride(0), L.ptr(attr), L.ptr(ew),
                                    ctx.graph.ptr(), N, E, ctx.training, g_out.data_ptr(), g_out.stride(0), L.ptr(dx), Cin, 0,
                                    dw.data_ptr(), db.data_ptr(), dg.data_ptr(), dbe.data_ptr(), tape.data_ptr(),
                                    ws.data_ptr(), ws.numel(), L.stream()), 'edge1_bwd')
        return None, None, None, g_out, dx, None, None, dw, (db if ctx.has_bias else None), dg, dbe


# --------------------------------------------------------------------------------------------------
# one [Linear, BatchNorm1d?, ReLU?] stage of gcn_lib.sparse.MLP   gcn_lib/sparse/torch_nn.py:50-71
# --------------------------------------------------------------------------------------------------
class MLPStageFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, flags, buffers, x, w, b, gamma, beta):
        L.require_cuda(x, w)
        lib = L.lib()
        x, w = L.f32c(x), L.f32c(w)
        b = L.f32c(b) if b is not None else None
        M, K = x.shape
        Nout = w.shape[0]
        rm, rv, nbt = buffers if buffers is not None else (None, None, None)
        bn = _bn_struct(gamma, beta, rm, rv, nbt) if flags & BN else None
        y = torch.empty(M, Nout, dtype=torch.float32, device=x.device)
        tape_n = lib.yolat_mlp_tape_floats(M, K, Nout, flags)
        tape = torch.empty(max(tape_n, 1), dtype=torch.float32, device=x.device)
        ws = L.workspace.get(lib.yolat_mlp_ws_floats(M, K, Nout, flags), x.device)
        L.check(lib.yolat_mlp_fwd(x.data_ptr(), x.stride(0), M, K, w.data_ptr(), L.ptr(b), Nout,
                                  C.byref(bn) if bn is not None else None, flags, y.data_ptr(), y.stride(0),
                                  tape.data_ptr(), tape.numel(), ws.data_ptr(), ws.numel(), L.stream()), 'mlp_fwd')
        ctx.flags, ctx.buffers, ctx.has_bias = flags, buffers, b is not None
        ctx.bias_ptr = b.data_ptr() if b is not None else 0
        ctx.save_for_backward(x, w, gamma, beta, tape)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = L.lib()
        x, w, gamma, beta, tape = ctx.saved_tensors
        flags = ctx.flags
        M, K = x.shape
        Nout = w.shape[0]
        gy = L.f32rows(gy)
        rm, rv, _ = ctx.buffers if ctx.buffers is not None else (None, None, None)
        bn = _bn_struct(gamma, beta, rm, rv, None) if flags & BN else None
        dx = torch.empty_like(x) if ctx.needs_input_grad[2] else None
        dw = _empty_like_param(w)
        db = None
        if ctx.has_bias:
            e = _GRAD_ARENA.get(ctx.bias_ptr)
            db = e[0].narrow(0, e[1], e[2]) if e is not None else torch.empty(Nout, dtype=torch.float32, device=x.device)
        dg = _empty_like_param(gamma) if flags & BN else None
        dbe = _empty_like_param(beta) if flags & BN else None
        ws = L.workspace.get(lib.yolat_mlp_ws_floats(M, K, Nout, flags), x.device)
        L.check(lib.yolat_mlp_bwd(x.data_ptr(), x.stride(0), M, K, w.data_ptr(), Nout,
                                  C.byref(bn) if bn is not None else None, flags, gy.data_ptr(), gy.stride(0),
                                  L.ptr(dx), K, 0, dw.data_ptr(), L.ptr(db), L.ptr(dg), L.ptr(dbe),
                                  tape.data_ptr(), ws.data_ptr(), ws.numel(), L.stream()), 'mlp_bwd')
        return None, None, dx, dw, db, dg, dbe


def mlp_stage(x, lin, bn=None, relu=False, training=True):
    """y = act(bn(lin(x))) through the fused stage kernel; `lin`/`bn` are nn.Linear / nn.BatchNorm1d."""
    flags = (BN if bn is not None else 0) | (RELU if relu else 0)
    use_batch_stats = bn is not None and (training or bn.running_mean is None)
    if use_batch_stats:
        flags |= TRAINING
    buffers = None
    gamma = beta = None
    if bn is not None:
        buffers = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if training else \
                  (bn.running_mean, bn.running_var, None)
        gamma, beta = bn.weight, bn.bias
    return MLPStageFn.apply(flags, buffers, x, lin.weight, lin.bias, gamma, beta)


# --------------------------------------------------------------------------------------------------
# torch_scatter.scatter(dim=0, reduce='mean'|'max')   architecture3cc_rpn_gp_iter2.py:67,122
# --------------------------------------------------------------------------------------------------
class SegmentMeanFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, seg, src):
        L.require_cuda(src)
        src = L.f32c(src)
        M, Cc = src.shape
        out = torch.empty(seg.S, Cc, dtype=torch.float32, device=src.device)
        L.check(L.lib().yolat_segment_mean_fwd(src.data_ptr(), src.stride(0), M, Cc, seg.ptr(), seg.S, out.data_ptr(),
                                               out.stride(0), L.stream()), 'segment_mean_fwd')
        ctx.seg, ctx.shape = seg, (M, Cc)
        return out

    @staticmethod
    def backward(ctx, g):
        g = L.f32rows(g)
        M, Cc = ctx.shape
        d = torch.zeros(M, Cc, dtype=torch.float32, device=g.device)   # rows with an out-of-range index keep 0
        L.check(L.lib().yolat_segment_mean_bwd(g.data_ptr(), g.stride(0), M, Cc, ctx.seg.ptr(), ctx.seg.S, d.data_ptr(),
                                               d.stride(0), 0, L.stream()), 'segment_mean_bwd')
        return None, d


class SegmentMaxFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, seg, src):
        L.require_cuda(src)
        src = L.f32c(src)
        M, Cc = src.shape
        out = torch.empty(seg.S, Cc, dtype=torch.float32, device=src.device)
        arg = torch.empty(seg.S, Cc, dtype=torch.int32, device=src.device)
        L.check(L.lib().yolat_segment_max_fwd(src.data_ptr(), src.stride(0), M, Cc, seg.ptr(), seg.S, out.data_ptr(),
                                              out.stride(0), arg.data_ptr(), L.stream()), 'segment_max_fwd')
        ctx.seg, ctx.shape = seg, (M, Cc)
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        g = L.f32rows(g)
        M, Cc = ctx.shape
        d = torch.empty(M, Cc, dtype=torch.float32, device=g.device)
        L.check(L.lib().yolat_segment_max_bwd(g.data_ptr(), g.stride(0), M, Cc, ctx.seg.S, arg.data_ptr(), d.data_ptr(),
                                              d.stride(0), 0, L.stream()), 'segment_max_bwd')
        return None, d


def segment_mean(src, seg):
    return SegmentMeanFn.apply(seg, src)


def segment_max(src, seg):
    return SegmentMaxFn.apply(seg, src)


# --------------------------------------------------------------------------------------------------
# fusion_block + cat + scatter-max   architecture3cc_rpn_gp_iter2.py:62-63,122
# --------------------------------------------------------------------------------------------------
class FuseMaxFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, seg, training, buffers, feats, w, b, gamma, beta):
        L.require_cuda(feats, w)
        lib = L.lib()
        feats, w = L.f32c(feats), L.f32c(w)
        b = L.f32c(b) if b is not None else None
        M, K = feats.shape
        F_ = w.shape[0]
        rm, rv, nbt = buffers
        bn = _bn_struct(gamma, beta, rm, rv, nbt if training else None)
        pooled = torch.empty(seg.S, F_ + K, dtype=torch.float32, device=feats.device)
        tape = torch.empty(max(lib.yolat_fusemax_tape_floats(M, K, F_, seg.S), 1), dtype=torch.float32,
                           device=feats.device)
        ws = L.workspace.get(lib.yolat_fusemax_ws_floats(M, K, F_, seg.S), feats.device)
        L.check(lib.yolat_fusemax_fwd(feats.data_ptr(), feats.stride(0), M, K, w.data_ptr(), L.ptr(b), F_, C.byref(bn),
                                      int(training), seg.ptr(), seg.S, pooled.data_ptr(), pooled.stride(0),
                                      tape.data_ptr(), tape.numel(), ws.data_ptr(), ws.numel(), L.stream()),
                'fusemax_fwd')
        ctx.seg, ctx.training, ctx.buffers, ctx.has_bias = seg, int(training), buffers, b is not None
        ctx.bias_ptr = b.data_ptr() if b is not None else 0
        ctx.save_for_backward(feats, w, gamma, beta, tape)
        return pooled

    @staticmethod
    def backward(ctx, gp):
        lib = L.lib()
        feats, w, gamma, beta, tape = ctx.saved_tensors
        seg = ctx.seg
        M, K = feats.shape
        F_ = w.shape[0]
        gp = L.f32rows(gp)
        rm, rv, _ = ctx.buffers
        bn = _bn_struct(gamma, beta, rm, rv, None)
        dfeats = torch.empty_like(feats) if ctx.needs_input_grad[3] else None
        dw = _empty_like_param(w)
        e = _GRAD_ARENA.get(ctx.bias_ptr) if ctx.has_bias else None
        db = e[0].narrow(0, e[1], e[2]) if e is not None else torch.empty(F_, dtype=torch.float32, device=w.device)
        dg, dbe = _empty_like_param(gamma), _empty_like_param(beta)
        ws = L.workspace.get(lib.yolat_fusemax_ws_floats(M, K, F_, seg.S), feats.device)
        L.check(lib.yolat_fusemax_bwd(feats.data_ptr(), feats.stride(0), M, K, w.data_ptr(), F_, C.byref(bn),
                                      ctx.training, seg.ptr(), seg.S, gp.data_ptr(), gp.stride(0), L.ptr(dfeats), K, 0,
                                      dw.data_ptr(), db.data_ptr(), dg.data_ptr(), dbe.data_ptr(), tape.data_ptr(),
                                      ws.data_ptr(), ws.numel(), L.stream()), 'fusemax_bwd')
        return None, None, None, dfeats, dw, (db if ctx.has_bias else None), dg, dbe


# --------------------------------------------------------------------------------------------------
# CrossEntropyLoss (mean)   architecture3cc_rpn_gp_iter2.py:363,376
# --------------------------------------------------------------------------------------------------
class SoftmaxXentFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, logits, labels):
        L.require_cuda(logits, labels)
        logits = L.f32c(logits)
        labels = labels.long().contiguous()
        B, ncls = logits.shape
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        prob = torch.empty(B, ncls, dtype=torch.float32, device=logits.device)
        ws = L.workspace.get(B, logits.device)
        L.check(L.lib().yolat_softmax_xent_fwd(logits.data_ptr(), logits.stride(0), B, ncls, labels.data_ptr(),
                                               loss.data_ptr(), prob.data_ptr(), ws.data_ptr(), ws.numel(),
                                               L.stream()), 'softmax_xent_fwd')
        ctx.save_for_backward(prob, labels)
        return loss

    @staticmethod
    def backward(ctx, g):
        prob, labels = ctx.saved_tensors
        B, ncls = prob.shape
        g = L.f32c(g)
        d = torch.empty_like(prob)
        L.check(L.lib().yolat_softmax_xent_bwd(prob.data_ptr(), B, ncls, labels.data_ptr(), g.data_ptr(), d.data_ptr(),
                                               d.stride(0), L.stream()), 'softmax_xent_bwd')
        return d, None


def softmax_cross_entropy(logits, labels):
    return SoftmaxXentFn.apply(logits, labels)


# --------------------------------------------------------------------------------------------------
# raw dense product on the tensor cores (the block under every nn.Linear; no autograd)
# --------------------------------------------------------------------------------------------------
GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2


def gemm(mode, a, b, bias=None, out=None, accumulate=False):
    """mode NT: a[M,K] @ b[N,K]^T;  NN: a[M,K] @ b[K,N];  TN: a[K,M]^T @ b[K,N]  (+ bias[N]); fp32, row-major,
    arbitrary row strides."""
    L.require_cuda(a, b)
    lib = L.lib()
    if a.dtype != torch.float32 or b.dtype != torch.float32 or a.stride(1) != 1 or b.stride(1) != 1:
        raise ValueError('gemm operands must be fp32 with unit inner stride')
    if mode == GEMM_NT:
        (M, K), N = a.shape, b.shape[0]
    elif mode == GEMM_NN:
        (M, K), N = a.shape, b.shape[1]
    else:
        (K, M), N = a.shape, b.shape[1]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    ws = L.workspace.get(lib.yolat_gemm_ws_floats(mode, M, N, K), a.device)
    L.check(lib.yolat_gemm(mode, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), out.stride(0),
                           M, N, K, L.ptr(bias), int(accumulate), ws.data_ptr(), ws.numel(), L.stream()), 'gemm')
    return out
