from .conv import MessagePassing
from . import data_parallel  # noqa: F401


class _Stub(MessagePassing):
    def __init__(self, *a, **k):
        raise NotImplementedError('PyG conv zoo is not on the YOLaT hot path')


class GATConv(_Stub):
    pass


class SAGEConv(_Stub):
    pass


class GCNConv(_Stub):
    pass


class GINConv(_Stub):
    pass
