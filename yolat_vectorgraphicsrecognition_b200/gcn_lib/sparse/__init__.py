"""Drop-in for `gcn_lib.sparse` (reference: gcn_lib/sparse/__init__.py:1-3 star-exports torch_nn,
torch_edge, torch_vertex).  The classes on the YOLaT hot path run on the sm_100a kernels; the DeepGCN
leftovers the reference never instantiates are importable names that raise on construction."""
from .torch_nn import MLP, MultiSeq, act_layer, norm_layer
from .torch_vertex import (AttrRelativeEdgeConvGlobalPool2, GraphConv, ResBlock, PlainDynBlock, DenseDynBlock,
                           ResDynBlock, DynConv, DilatedKnnGraph, ResGraphBlock, DenseGraphBlock)

__all__ = ['MLP', 'MultiSeq', 'act_layer', 'norm_layer', 'AttrRelativeEdgeConvGlobalPool2', 'GraphConv', 'ResBlock',
           'PlainDynBlock', 'DenseDynBlock', 'ResDynBlock', 'DynConv', 'DilatedKnnGraph', 'ResGraphBlock',
           'DenseGraphBlock']
