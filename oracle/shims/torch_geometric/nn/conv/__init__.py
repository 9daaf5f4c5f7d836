import inspect
import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    """Restatement of PyG 1.6 MessagePassing for flow='source_to_target', node_dim=0."""

    def __init__(self, aggr='add', flow='source_to_target', node_dim=0, **kwargs):
        super().__init__()
        assert flow == 'source_to_target'
        self.aggr = aggr
        self.node_dim = node_dim
        self._msg_params = list(inspect.signature(self.message).parameters.keys())

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        N = None
        args = {}
        for name in self._msg_params:
            if name.endswith('_i') or name.endswith('_j'):
                base = name[:-2]
                val = kwargs[base]
                sel = i if name.endswith('_i') else j
                if isinstance(val, (tuple, list)):
                    val = val[1] if name.endswith('_i') else val[0]
                if val is None:
                    args[name] = None
                else:
                    if name.endswith('_i'):
                        N = val.shape[0]
                    args[name] = val.index_select(0, sel)
            else:
                args[name] = kwargs.get(name)
        if N is None:
            x = kwargs.get('x')
            x = x[1] if isinstance(x, (tuple, list)) else x
            N = x.shape[0]
        msg = self.message(**args)
        return scatter(msg, i, dim=0, dim_size=N, reduce=self.aggr)

    def message(self, x_j):
        return x_j
