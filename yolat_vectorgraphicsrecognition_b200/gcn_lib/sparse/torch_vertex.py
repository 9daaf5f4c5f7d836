"""`GraphConv`, `ResBlock` and the `attr_edge_gp2` conv with the reference's signatures, attribute paths
and state-dict keys (gcn_lib/sparse/torch_vertex.py:288-341, 730-775, 808-829), running on the
sm_100a kernels behind include/yolat_b200.h.

Only `conv='attr_edge_gp2'` is on YOLaT's hot path: `Backbone.__init__` hard-codes it
(cad_recognition/architecture3cc_rpn_gp_iter2.py:22).  The sibling recipes of the same family -- 'edge', 'attr_edge',
'multilayer_edge', 'attr_edge_gp' (torch_vertex.py:738-747) -- run on the same kernels: their message MLP reads a
sub-set / permutation of [x_i | x_j - x_i | attr], so the recipe's first Linear weight is embedded (a differentiable
torch op on a [C, F] tensor) into that column layout and the node-level P | Q trick applies unchanged.
They need act='relu' and norm='batch' (what YOLaT passes, architecture...py:19-20); the remaining DeepGCN convs raise
NotImplementedError with the reference's message format.
"""
import torch
from torch import nn

from ... import ops
from ...graph import graph_for
from .torch_nn import MLP, MultiSeq


def reset(module):
    """torch_vertex.py:111-121 -- re-initialise the children (keeps the reference's RNG consumption, so
    `torch.manual_seed(s)` followed by construction yields the same parameters as the reference)."""
    def _reset(item):
        if hasattr(item, 'reset_parameters'):
            item.reset_parameters()
    if module is not None:
        if hasattr(module, 'children') and len(list(module.children())) > 0:
            for item in module.children():
                _reset(item)
        else:
            _reset(module)


class AttrRelativeEdgeConvGlobalPool2(nn.Module):
    """torch_vertex.py:288-341.  message = nn([x_i, x_j - x_i, attr]) (* norm), aggr = 'mean' at the
    target, `out += lin_r(x)`, `x_node = mlp_node(x_node)`; returns (out, x_node)."""

    def __init__(self, in_channels, out_channels, **kwargs):
        super(AttrRelativeEdgeConvGlobalPool2, self).__init__()
        self.aggr = 'mean'
        self.nn = MLP([in_channels * 2 + 4, out_channels, out_channels], 'relu', 'batch')
        self.lin_r = torch.nn.Linear(in_channels, out_channels, bias=True)
        self.mlp_node = MLP([in_channels, out_channels], 'relu', 'batch')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.reset_parameters()

    def reset_parameters(self):
        reset(self.nn)

    def _flat(self):
        l1, b1, _, l2, b2, _ = list(self.nn.children())
        ln, bn_ = list(self.mlp_node.children())[:2]
        params = (l1.weight, l1.bias, b1.weight, b1.bias, l2.weight, l2.bias, b2.weight, b2.bias,
                  self.lin_r.weight, self.lin_r.bias, ln.weight, ln.bias, bn_.weight, bn_.bias)
        buffers = (b1.running_mean, b1.running_var, b1.num_batches_tracked,
                   b2.running_mean, b2.running_var, b2.num_batches_tracked,
                   bn_.running_mean, bn_.running_var, bn_.num_batches_tracked)
        return params, buffers

    def forward(self, x, x_node, edge_index, edge_weight=None, edge_attr=None):
        if isinstance(x, (tuple, list)):          # PairTensor: the reference uses x[1] for lin_r and x_i
            if x[0] is not x[1]:
                raise NotImplementedError('attr_edge_gp2: bipartite (x_src, x_dst) inputs are not supported')
            x = x[1]
        graph = graph_for(edge_index, x.shape[0])
        params, buffers = self._flat()
        return ops.GP2ConvFn.apply(graph, self.training, torch.is_grad_enabled(), buffers, x, x_node, edge_attr,
                                   edge_weight, *params)

    def __repr__(self):
        return '{}(nn={})'.format(self.__class__.__name__, self.nn)


class _EdgeRecipe(nn.Module):
    """Common part of the one- and two-stage edge convolutions: `self.nn` (the message MLP), `self.lin_r`, the unused
    `self.mlp` / `self.lin_l` children the reference constructs (they are part of its state dict), and the embedding of
    nn.0.weight into the [x_i | x_j - x_i | attr] layout."""

    def _require(self, act, norm, name):
        if (act or '').lower() != 'relu' or (norm or '').lower() != 'batch':
            raise NotImplementedError('conv {} needs act=\'relu\' and norm=\'batch\' in yolat_b200 (got act={}, norm={})'
                                      .format(name, act, norm))

    def reset_parameters(self):
        reset(self.nn)

    def _embed(self, w, Cin):
        """recipe weight [C, F] -> [C, 2 Cin + 4]; `self.parts` lists the recipe's concat order with entries
        'i' (x_i), 'd' (x_j - x_i), 'a' (attr)."""
        cols, off = {}, 0
        for p in self.parts:
            n = 4 if p == 'a' else Cin
            cols[p] = w[:, off:off + n]
            off += n
        z = w.new_zeros(w.shape[0], Cin)
        return torch.cat([cols.get('i', z), cols.get('d', z), cols.get('a', w.new_zeros(w.shape[0], 4))], dim=1)

    def _x(self, x):
        if isinstance(x, (tuple, list)):
            if x[0] is not x[1]:
                raise NotImplementedError('bipartite (x_src, x_dst) inputs are not supported')
            x = x[1]
        return x

    def _attr(self, graph, edge_attr, x):
        if edge_attr is None:
            return x.new_zeros(graph.E, 4)
        return edge_attr

    def _one_stage(self, base, x, edge_index, edge_weight, edge_attr, Cin):
        graph = graph_for(edge_index, x.shape[0])
        lin, bn = list(self.nn.children())[:2]
        return ops.Edge1ConvFn.apply(graph, self.training, (bn.running_mean, bn.running_var, bn.num_batches_tracked), base,
                                     x, self._attr(graph, edge_attr, x), edge_weight, self._embed(lin.weight, Cin), lin.bias,
                                     bn.weight, bn.bias)

    def __repr__(self):
        return '{}(nn={})'.format(self.__class__.__name__, self.nn)


class EdgConv(_EdgeRecipe):
    """conv='edge': WeightedRelativeEdgeConv (torch_vertex.py:427-484) with nn = MLP([2 Cin, C]) (:546-557):
    message = nn([x_j - x_i, x_i]) (* norm), mean aggregation, out += lin_r(x)."""
    parts = ('d', 'i')

    def __init__(self, in_channels, out_channels, act='relu', norm=None, bias=True, aggr='add'):
        super(EdgConv, self).__init__()
        self._require(act, norm, 'edge')
        self.nn = MLP([in_channels * 2, out_channels], act, norm, bias)
        self.mlp = MultiSeq(*[MLP([in_channels, 64]), MLP([64, in_channels])])          # unused by forward (:449-451)
        self.lin_l = torch.nn.Linear(out_channels, out_channels, bias=True)            # unused by forward (:453)
        self.lin_r = torch.nn.Linear(in_channels, out_channels, bias=True)
        self.in_channels = in_channels
        self.reset_parameters()

    def forward(self, x, edge_index, edge_weight=None):
        x = self._x(x)
        return self._one_stage(self.lin_r(x), x, edge_index, edge_weight, None, self.in_channels)


class AttrEdgConv(_EdgeRecipe):
    """conv='attr_edge': AttrRelativeEdgeConv (torch_vertex.py:219-286) with nn = MLP([Cin + 4, C]) (:560-573):
    message = nn([x_j - x_i, attr]) (* norm)."""
    parts = ('d', 'a')

    def __init__(self, in_channels, out_channels, act='relu', norm=None, bias=True, aggr='add'):
        super(AttrEdgConv, self).__init__()
        self._require(act, norm, 'attr_edge')
        self.nn = self._make_nn(in_channels, out_channels, act, norm, bias)
        self.mlp = MultiSeq(*[MLP([in_channels, 64]), MLP([64, in_channels])])          # unused by forward (:240-242)
        self.lin_r = torch.nn.Linear(in_channels, out_channels, bias=True)
        self.in_channels = in_channels
        self.reset_parameters()

    def _make_nn(self, cin, cout, act, norm, bias):
        return MLP([cin + 4, cout], act, norm, bias)

    def forward(self, x, edge_index, edge_weight=None, edge_attr=None):
        x = self._x(x)
        return self._one_stage(self.lin_r(x), x, edge_index, edge_weight, edge_attr, self.in_channels)


class MultilayerEdgConv(AttrEdgConv):
    """conv='multilayer_edge' (torch_vertex.py:593-605): the same message with a TWO-stage MLP([Cin + 4, C, C]) -- exactly
    the edge branch of attr_edge_gp2 with a zero x_i block, so it runs on the fused K-EDGE kernels (yolat_gp2_*)."""

    def _make_nn(self, cin, cout, act, norm, bias):
        return MLP([cin + 4, cout, cout], act, norm, bias)

    def forward(self, x, edge_index, edge_weight=None, edge_attr=None):
        x = self._x(x)
        graph = graph_for(edge_index, x.shape[0])
        l1, b1, _, l2, b2, _ = list(self.nn.children())
        C_ = l2.weight.shape[0]
        if not hasattr(self, '_dummy') or self._dummy[0].device != x.device:
            # the node branch of the GP2 entry point is not part of this recipe: constant dummy parameters, output dropped
            self._dummy = (torch.zeros(C_, self.in_channels, device=x.device), torch.zeros(C_, device=x.device),
                           torch.ones(C_, device=x.device), torch.zeros(C_, device=x.device),
                           torch.zeros(C_, device=x.device), torch.ones(C_, device=x.device))
        wn, bnode, gn, ben, rmn, rvn = self._dummy
        params = (self._embed(l1.weight, self.in_channels), l1.bias, b1.weight, b1.bias, l2.weight, l2.bias, b2.weight,
                  b2.bias, self.lin_r.weight, self.lin_r.bias, wn, bnode, gn, ben)
        buffers = (b1.running_mean, b1.running_var, b1.num_batches_tracked, b2.running_mean, b2.running_var,
                   b2.num_batches_tracked, rmn, rvn, None)
        out, _ = ops.GP2ConvFn.apply(graph, self.training, torch.is_grad_enabled(), buffers, x, x.detach(),
                                     self._attr(graph, edge_attr, x), edge_weight, *params)
        return out


class EdgConvGlobalPool(_EdgeRecipe):
    """conv='attr_edge_gp': AttrRelativeEdgeConvGlobalPool (torch_vertex.py:343-425) with nn = MLP([2 Cin + 4, C])
    (:575-590).  x carries [features | root features] (2 Cin columns): message = nn([x_i, x_j - x_i, attr]) on the first
    Cin columns, out += lin_r(x[:, :Cin]) + mlp(x[:, Cin:])."""
    parts = ('i', 'd', 'a')

    def __init__(self, in_channels, out_channels, act='relu', norm=None, bias=True, aggr='add'):
        super(EdgConvGlobalPool, self).__init__()
        self._require(act, norm, 'attr_edge_gp')
        self.nn = MLP([in_channels * 2 + 4, out_channels], act, norm, bias)
        self.mlp = MultiSeq(*[MLP([in_channels, out_channels])])
        self.pool = torch.nn.AdaptiveAvgPool1d(1)
        self.lin_r = torch.nn.Linear(in_channels, out_channels, bias=True)
        self.in_channels = in_channels
        self.reset_parameters()

    def forward(self, x, edge_index, edge_weight=None, edge_attr=None):
        x = self._x(x)
        xf = x[:, 0:self.in_channels].contiguous()
        base = self.lin_r(xf) + self.mlp(x[:, self.in_channels:].contiguous())
        return self._one_stage(base, xf, edge_index, edge_weight, edge_attr, self.in_channels)


_KNOWN_CONVS = ('attr_edge_cf', 'mr', 'gat', 'gcn', 'gin', 'sage', 'rsage')


class GraphConv(nn.Module):
    """Static graph convolution layer (torch_vertex.py:730-775)."""

    def __init__(self, in_channels, out_channels, conv='gcn', act='relu', norm=None, bias=True, heads=8):
        super(GraphConv, self).__init__()
        self.conv = conv.lower()
        if self.conv == 'attr_edge_gp2':
            # act / norm / bias are ignored by this conv in the reference too (torch_vertex.py:749)
            self.gconv = AttrRelativeEdgeConvGlobalPool2(in_channels, out_channels)
        elif self.conv == 'edge':
            self.gconv = EdgConv(in_channels, out_channels, act, norm, bias)
        elif self.conv == 'multilayer_edge':
            self.gconv = MultilayerEdgConv(in_channels, out_channels, act, norm, bias)
        elif self.conv == 'attr_edge':
            self.gconv = AttrEdgConv(in_channels, out_channels, act, norm, bias)
        elif self.conv == 'attr_edge_gp':
            self.gconv = EdgConvGlobalPool(in_channels, out_channels, act, norm, bias)
        elif self.conv in _KNOWN_CONVS:
            raise NotImplementedError('conv {} is not implemented'.format(conv) +
                                      ' by yolat_b200 (only attr_edge_gp2 is on the YOLaT hot path)')
        else:
            raise NotImplementedError('conv {} is not implemented'.format(conv))

    def forward(self, x, edge_index, edge_weight=None, edge_attr=None, pos=None, x_node=None):
        # the dispatch of torch_vertex.py:765-775
        if self.conv in ('attr_edge', 'multilayer_edge', 'attr_edge_gp'):
            return self.gconv(x, edge_index, edge_weight, edge_attr)
        if self.conv == 'edge':
            return self.gconv(x, edge_index, edge_weight) if edge_weight is not None else self.gconv(x, edge_index)
        return self.gconv(x, x_node, edge_index, edge_weight, edge_attr)


class ResBlock(nn.Module):
    """Residual graph convolution block (torch_vertex.py:808-829).  For attr_edge_gp2 the reference adds
    NO residual (both `+=` lines are commented out, :825-826)."""

    def __init__(self, channels, conv='edge', act='relu', norm=None, bias=True, res_scale=1, **kwargs):
        super(ResBlock, self).__init__()
        self.body = GraphConv(channels, channels, conv, act, norm, bias, **kwargs)
        self.res_scale = res_scale
        self.channels = channels

    def forward(self, x, edge, edge_weight=None, edge_attr=None, pos=None, x_node=None):
        out, out_node = self.body(x, edge, edge_weight, edge_attr, x_node=x_node)
        return out, out_node


def _not_on_path(name):
    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError('%s is a DeepGCN leftover that YOLaT never instantiates; '
                                      'it is outside the yolat_b200 hot path' % name)
    _Stub.__name__ = name
    return _Stub


PlainDynBlock = _not_on_path('PlainDynBlock')
DenseDynBlock = _not_on_path('DenseDynBlock')
ResDynBlock = _not_on_path('ResDynBlock')
DynConv = _not_on_path('DynConv')
DilatedKnnGraph = _not_on_path('DilatedKnnGraph')
ResGraphBlock = _not_on_path('ResGraphBlock')
DenseGraphBlock = _not_on_path('DenseGraphBlock')
