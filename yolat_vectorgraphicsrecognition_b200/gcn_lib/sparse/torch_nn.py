"""`MLP`, `MultiSeq`, `act_layer`, `norm_layer` with the reference's signatures and child naming
(gcn_lib/sparse/torch_nn.py:9-71), executing [Linear, BatchNorm1d, ReLU] runs as fused stage kernels."""
from torch import nn
from torch.nn import Sequential as Seq, Linear as Lin

from ... import ops


def act_layer(act_type, inplace=False, neg_slope=0.2, n_prelu=1):
    """torch_nn.py:9-20"""
    act = act_type.lower()
    if act == 'relu':
        return nn.ReLU(inplace)
    if act == 'leakyrelu':
        return nn.LeakyReLU(neg_slope, inplace)
    if act == 'prelu':
        return nn.PReLU(num_parameters=n_prelu, init=neg_slope)
    raise NotImplementedError('activation layer [%s] is not found' % act)


def norm_layer(norm_type, nc):
    """torch_nn.py:23-34"""
    norm = norm_type.lower()
    if norm == 'batch':
        return nn.BatchNorm1d(nc, affine=True)
    if norm == 'layer':
        return nn.LayerNorm(nc, elementwise_affine=True)
    if norm == 'instance':
        return nn.InstanceNorm1d(nc, affine=False)
    raise NotImplementedError('normalization layer [%s] is not found' % norm)


class MultiSeq(Seq):
    """Sequential that splats tuple outputs into the next module (torch_nn.py:37-47)."""

    def __init__(self, *args):
        super(MultiSeq, self).__init__(*args)

    def forward(self, *inputs):
        for module in self._modules.values():
            if type(inputs) == tuple:
                inputs = module(*inputs)
            else:
                inputs = module(inputs)
        return inputs


def _fusable_bn(m):
    return (isinstance(m, nn.BatchNorm1d) and m.affine and m.track_running_stats and m.momentum == 0.1
            and m.eps == 1e-5)


class MLP(Seq):
    """[Lin, norm?, act?, Dropout2d?] per layer, integer child names (torch_nn.py:50-71)."""

    def __init__(self, channels, act='relu', norm=None, bias=True, drop=0., last_lin=False):
        m = []
        for i in range(1, len(channels)):
            m.append(Lin(channels[i - 1], channels[i], bias))
            if (i == len(channels) - 1) and last_lin:
                pass
            else:
                if norm is not None and norm.lower() != 'none':
                    m.append(norm_layer(norm, channels[i]))
                if act is not None and act.lower() != 'none':
                    m.append(act_layer(act))
                if drop > 0:
                    m.append(nn.Dropout2d(drop))
        self.m = m
        super(MLP, self).__init__(*self.m)

    def stages(self):
        """Children grouped as (Linear, BatchNorm1d | None, has_relu) runs + pass-through modules."""
        mods = list(self._modules.values())
        out, i = [], 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, Lin):
                bn, relu, j = None, False, i + 1
                if j < len(mods) and _fusable_bn(mods[j]):
                    bn, j = mods[j], j + 1
                if j < len(mods) and type(mods[j]) is nn.ReLU:
                    relu, j = True, j + 1
                out.append(('stage', m, bn, relu))
                i = j
            else:
                out.append(('module', m))
                i += 1
        return out

    def forward(self, x):
        for st in self.stages():
            if st[0] == 'stage':
                _, lin, bn, relu = st
                x = ops.mlp_stage(x, lin, bn, relu, training=self.training if bn is None else bn.training)
            else:
                x = st[1](x)
        return x
