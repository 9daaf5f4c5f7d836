"""yolat_vectorgraphicsrecognition_b200 -- B200-native (sm_100a) engine for YOLaT's Bezier-graph
proposal classifier: the `gcn_lib.sparse` GraphConv('attr_edge_gp2') / MLP layers, the
`torch_scatter.scatter` proposal pooling and `architecture3cc_rpn_gp_iter2.SparseCADGCN`, behind the
reference's own Python surface.  The compute path is libyolat_b200.so (include/yolat_b200.h)."""
from . import _lib  # noqa: F401

__version__ = '0.1.0'
