"""CPU restatement of the reference hot path, as plain functional torch (fp32 or fp64).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function cites the reference lines it follows
(paths relative to /root/reference).  It keeps the reference's op chain (index_select, cat, addmm,
batch_norm, relu, scatter) so it also serves as the "port" CPU baseline in bench.py.  Pinned against
the unmodified reference by oracle/make_golden.py -> tests/golden/*.pt (tests/test_oracle.py).

State is a flat dict with the reference's state-dict keys, e.g.
`cls_net.head.gconv.nn.0.weight`.  Parameters that require grad are leaf tensors in that dict.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm1d default, gcn_lib/sparse/torch_nn.py:27
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------
# third-party semantics restated (torch-scatter 2.0.x, PyG 1.6/1.7 -- not under /root/reference)
# ----------------------------------------------------------------------------------------------
def scatter_mean(src, index, dim_size):
    """torch_scatter.scatter(reduce='mean') over dim 0: sum / clamp(count, 1); empty segments -> 0."""
    out = src.new_zeros((dim_size,) + src.shape[1:]).index_add_(0, index, src)
    cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
    return out / cnt.view(-1, *([1] * (src.dim() - 1)))


class _ScatterMaxFn(torch.autograd.Function):
    """torch_scatter.scatter(reduce='max') over dim 0: value of the arg-max row, empty -> 0, the
    gradient goes to a single arg index (first occurrence, as the torch_scatter CPU kernel)."""

    @staticmethod
    def forward(ctx, src, index, dim_size):
        M, Fd = src.shape
        idx2 = index.view(-1, 1).expand(M, Fd)
        out = src.new_full((dim_size, Fd), float('-inf')).scatter_reduce(0, idx2, src, 'amax', include_self=True)
        rows = torch.arange(M).view(-1, 1).expand(M, Fd)
        cand = torch.where(src == out.gather(0, idx2), rows, torch.full_like(rows, M))
        arg = torch.full((dim_size, Fd), M, dtype=torch.long).scatter_reduce(0, idx2, cand, 'amin', include_self=True)
        empty = arg == M
        ctx.save_for_backward(arg, empty)
        ctx.M = M
        return torch.where(empty, torch.zeros_like(out), out)

    @staticmethod
    def backward(ctx, g):
        arg, empty = ctx.saved_tensors
        g = torch.where(empty, torch.zeros_like(g), g)
        return g.new_zeros((ctx.M + 1, g.shape[1])).scatter_(0, arg, g)[:ctx.M], None, None


def scatter_max(src, index, dim_size):
    return _ScatterMaxFn.apply(src, index, dim_size)


# ----------------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------------
def _bn(state, prefix, z, training):
    """nn.BatchNorm1d(nc, affine=True): batch stats (biased var) in training, running stats in eval;
    running buffers updated with momentum 0.1 and the UNBIASED variance (torch_nn.py:27)."""
    rm, rv = state[prefix + '.running_mean'], state[prefix + '.running_var']
    y = F.batch_norm(z, rm, rv, state[prefix + '.weight'], state[prefix + '.bias'], training, BN_MOMENTUM, BN_EPS)
    if training and (prefix + '.num_batches_tracked') in state:
        state[prefix + '.num_batches_tracked'] += 1
    return y


def mlp(state, prefix, x, n_layers, training, norm=True, act=True):
    """gcn_lib.sparse.MLP (torch_nn.py:50-71): [Lin, BN, ReLU] * n_layers with Sequential child indices."""
    stride = 1 + int(norm) + int(act)
    for l in range(n_layers):
        base = '%s.%d' % (prefix, l * stride) if prefix else str(l * stride)
        x = F.linear(x, state[base + '.weight'], state.get(base + '.bias'))
        if norm:
            x = _bn(state, '%s.%d' % (prefix, l * stride + 1), x, training)
        if act:
            x = F.relu(x)
    return x


def gp2_conv(state, prefix, x, x_node, edge_index, attr, training, edge_weight=None):
    """AttrRelativeEdgeConvGlobalPool2 (gcn_lib/sparse/torch_vertex.py:288-341).

    propagate (PyG, flow source_to_target): x_j = x[edge_index[0]], x_i = x[edge_index[1]];
    message :330-337; aggr='mean' at the target :308; `out += lin_r(x)` :325; `mlp_node(x_node)` :326.
    """
    j, i = edge_index[0], edge_index[1]
    x_i = x.index_select(0, i)
    x_j = x.index_select(0, j)
    f = torch.cat([x_i, x_j - x_i, attr], dim=1)                       # :331
    m = mlp(state, prefix + '.nn', f, 2, training)                      # :309, :335
    if edge_weight is not None:
        m = edge_weight.view(-1, 1) * m                                 # :337
    out = scatter_mean(m, i, x.shape[0])                                # aggr='mean'
    out = out + F.linear(x, state[prefix + '.lin_r.weight'], state[prefix + '.lin_r.bias'])   # :325
    xn = mlp(state, prefix + '.mlp_node', x_node, 1, training)          # :326
    return out, xn


def backbone(state, opt, x, edge_index, attr, bbox_idx, n_prop, training, prefix='cls_net'):
    """Backbone.forward (cad_recognition/architecture3cc_rpn_gp_iter2.py:44-71)."""
    f, fs = gp2_conv(state, prefix + '.head.gconv', x, x, edge_index, attr, training)            # :45
    feats, feats_super = [f], [fs]
    for b in range(opt.n_blocks - 1):                                                            # :50-57
        f, fs = gp2_conv(state, '%s.backbone.%d.body.gconv' % (prefix, b), feats[-1], feats_super[-1],
                         edge_index, attr, training)        # ResBlock adds no residual, torch_vertex.py:823-827
        feats.append(f)
        feats_super.append(fs)
    sel = range(opt.n_blocks - opt.n_blocks_out, opt.n_blocks)
    feats = torch.cat([feats[k] for k in sel], dim=1)                                            # :60-61
    fusion = mlp(state, prefix + '.fusion_block', feats, 1, training)                            # :62
    out_feat = torch.cat((fusion, feats), dim=1)                                                 # :63
    feats_super = torch.cat([feats_super[k] for k in sel], dim=1)                                # :65-66
    feats_super = scatter_mean(feats_super, bbox_idx, n_prop)                                    # :67
    fusion_super = mlp(state, prefix + '.fusion_block_super', feats_super, 1, training)          # :68
    out_feat_super = torch.cat((fusion_super, feats_super), dim=1)                               # :69
    return out_feat, out_feat_super


def cadgcn_forward(state, opt, x, edge, e_attr, bbox_idx, training=True):
    """SparseCADGCN.forward (architecture3cc_rpn_gp_iter2.py:106-137); returns logits [B, ncls]."""
    edge_index = edge.t()                                                                        # :110
    n_prop = int(bbox_idx.max()) + 1                      # scatter's implicit dim_size
    out_feat, out_super = backbone(state, opt, x, edge_index, e_attr, bbox_idx, n_prop, training)
    pooled = scatter_max(out_feat, bbox_idx, n_prop)                                             # :122
    h = torch.cat([pooled, out_super], dim=1)                                                    # :127
    h = mlp(state, 'prediction_cls.0', h, 1, training)                                           # :91
    h = mlp(state, 'prediction_cls.1', h, 1, training)                                           # :92 (dropout 0)
    logits = mlp(state, 'prediction_cls.2', h, 1, training, norm=False, act=False)               # :93
    if opt.classifier != 'softmax':
        logits = torch.sigmoid(logits)                                                           # :132-133
    return logits


def detection_loss(opt, logits, labels):
    """DetectionLoss.forward (architecture3cc_rpn_gp_iter2.py:368-379)."""
    if opt.classifier == 'softmax':
        return F.cross_entropy(logits, labels)
    onehot = torch.zeros_like(logits).scatter_(1, labels.unsqueeze(1), 1)
    return F.binary_cross_entropy(logits, onehot)


# ----------------------------------------------------------------------------------------------
# state construction
# ----------------------------------------------------------------------------------------------
def state_spec(opt):
    """(key, shape, kind) for every state-dict entry in the reference's construction order (Appendix A.3)."""
    C, Cin, nc = opt.n_filters, opt.in_channels, opt.n_classes
    spec = []

    def lin(p, o, i):
        spec.append((p + '.weight', (o, i), 'lin_w'))
        spec.append((p + '.bias', (o,), 'lin_b'))

    def bn(p, c):
        spec.extend([(p + '.weight', (c,), 'bn_w'), (p + '.bias', (c,), 'bn_b'),
                     (p + '.running_mean', (c,), 'bn_rm'), (p + '.running_var', (c,), 'bn_rv'),
                     (p + '.num_batches_tracked', (), 'bn_nbt')])

    def conv(p, ci):
        lin(p + '.nn.0', C, 2 * ci + 4); bn(p + '.nn.1', C)
        lin(p + '.nn.3', C, C); bn(p + '.nn.4', C)
        lin(p + '.lin_r', C, ci)
        lin(p + '.mlp_node.0', C, ci); bn(p + '.mlp_node.1', C)

    conv('cls_net.head.gconv', Cin)
    for b in range(opt.n_blocks - 1):
        conv('cls_net.backbone.%d.body.gconv' % b, C)
    fd = C + C * (opt.n_blocks_out - 1)
    lin('cls_net.fusion_block.0', 1024, fd); bn('cls_net.fusion_block.1', 1024)
    lin('cls_net.fusion_block_super.0', 1024, fd); bn('cls_net.fusion_block_super.1', 1024)
    lin('prediction_cls.0.0', 512, (fd + 1024) * 2); bn('prediction_cls.0.1', 512)
    lin('prediction_cls.1.0', 256, 512); bn('prediction_cls.1.1', 256)
    lin('prediction_cls.2.0', nc, 256)
    return spec


def clone_state(sd, dtype=torch.float32, requires_grad=True):
    """Deep-copies a state dict into leaf tensors of `dtype` (integer buffers keep their dtype)."""
    out = {}
    for k, v in sd.items():
        v = v.detach().cpu().clone()
        if v.is_floating_point():
            v = v.to(dtype)
            if requires_grad and not ('running_' in k):
                v.requires_grad_(True)
        out[k] = v
    return out


def run_step(state, opt, batch, training=True):
    """fwd + loss + bwd; returns dict(logits, loss, grads{key: tensor})."""
    logits = cadgcn_forward(state, opt, batch.x.to(_dt(state)), batch.edge, batch.e_attr.to(_dt(state)),
                            batch.bbox_idx, training)
    loss = detection_loss(opt, logits, batch.labels)
    params = {k: v for k, v in state.items() if v.requires_grad}
    grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
    return dict(logits=logits.detach(), loss=loss.detach(),
                grads={k: g for k, g in zip(params.keys(), grads)})


def _dt(state):
    return next(v.dtype for v in state.values() if v.is_floating_point())
