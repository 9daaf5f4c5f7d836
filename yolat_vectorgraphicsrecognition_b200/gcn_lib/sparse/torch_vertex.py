"""`GraphConv`, `ResBlock` and the `attr_edge_gp2` conv with the reference's signatures, attribute paths
and state-dict keys (gcn_lib/sparse/torch_vertex.py:288-341, 730-775, 808-829), running on the
sm_100a kernels behind include/yolat_b200.h.

Only `conv='attr_edge_gp2'` is on YOLaT's hot path: `Backbone.__init__` hard-codes it
(cad_recognition/architecture3cc_rpn_gp_iter2.py:22).  The other conv names the reference's dispatcher
accepts are not built here and raise NotImplementedError with the reference's message format.
"""
import torch
from torch import nn

from ... import ops
from ...graph import graph_for
from .torch_nn import MLP


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


_KNOWN_CONVS = ('edge', 'multilayer_edge', 'attr_edge', 'attr_edge_cf', 'attr_edge_gp', 'mr', 'gat', 'gcn', 'gin',
                'sage', 'rsage')


class GraphConv(nn.Module):
    """Static graph convolution layer (torch_vertex.py:730-775)."""

    def __init__(self, in_channels, out_channels, conv='gcn', act='relu', norm=None, bias=True, heads=8):
        super(GraphConv, self).__init__()
        self.conv = conv.lower()
        if self.conv == 'attr_edge_gp2':
            # act / norm / bias are ignored by this conv in the reference too (torch_vertex.py:749)
            self.gconv = AttrRelativeEdgeConvGlobalPool2(in_channels, out_channels)
        elif self.conv in _KNOWN_CONVS:
            raise NotImplementedError('conv {} is not implemented'.format(conv) +
                                      ' by yolat_b200 (only attr_edge_gp2 is on the YOLaT hot path)')
        else:
            raise NotImplementedError('conv {} is not implemented'.format(conv))

    def forward(self, x, edge_index, edge_weight=None, edge_attr=None, pos=None, x_node=None):
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
