"""Once-per-batch graph preparation (host side of csrc/graph.cu).

`CSRGraph` replaces the gather/scatter bookkeeping PyG's `MessagePassing.propagate` redoes on every
call (gcn_lib/sparse/torch_vertex.py:324): it is built once from `edge_index` and shared by every
GraphConv layer of the step and by backward.  `Segments` does the same for `bbox_idx`
(torch_scatter.scatter at cad_recognition/architecture3cc_rpn_gp_iter2.py:67,122).
"""
import torch

from . import _lib as L


class CSRGraph(object):
    """CSR-by-target + CSR-by-source of `edge_index` ([2,E] int64, row 0 = source j, row 1 = target i)."""

    def __init__(self, edge_index, num_nodes, check=False):
        L.require_cuda(edge_index)
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError('edge_index must be [2, E], got %s' % (tuple(edge_index.shape),))
        if edge_index.dtype != torch.int64:
            edge_index = edge_index.long()
        self.N = int(num_nodes)
        self.E = int(edge_index.shape[1])
        self.device = edge_index.device
        self._keep = edge_index            # pins the storage while the graph is alive
        lib = L.lib()
        n_ints = lib.yolat_graph_ints(self.N, self.E)
        self.buf = torch.empty(max(int(n_ints), 1), dtype=torch.int32, device=self.device)
        # element (e, c) lives at edge[e*stride_e + c*stride_c]; works for the transposed [E,2] view as is
        L.check(lib.yolat_graph_build(edge_index.data_ptr(), edge_index.stride(1), edge_index.stride(0), self.E,
                                      self.N, self.buf.data_ptr(), L.stream()), 'graph_build')
        if check and self.errors() != 0:
            raise L.YolatError('edge_index has %d entries outside [0, %d)' % (self.errors(), self.N))

    def ptr(self):
        return self.buf.data_ptr()

    def _view(self, fn, n):
        lib = L.lib()
        p = fn(self.buf.data_ptr(), self.N, self.E)
        off = (p - self.buf.data_ptr()) // 4
        return self.buf[off:off + n]

    def errors(self):
        """Number of out-of-range edges dropped by the build (host sync)."""
        return int(self._view(L.lib().yolat_graph_error_ptr, 1).item())

    def rowptr(self):
        return self._view(L.lib().yolat_graph_rowptr, self.N + 1)

    def src(self):
        return self._view(L.lib().yolat_graph_src, self.E)

    def eid(self):
        return self._view(L.lib().yolat_graph_eid, self.E)


class Segments(object):
    """Rows grouped by an index vector (proposals from `bbox_idx`); `num_segments` = scatter's dim_size."""

    def __init__(self, index, num_segments=None, side_stream=None):
        """`side_stream`: build on that stream, forked off the current one (stream-ordered, capturable); the caller must
        `join()` before the first use -- the proposal index is only needed after the GraphConv stack, so its four small
        kernels leave the critical path."""
        L.require_cuda(index)
        if index.dtype != torch.int64:
            index = index.long()
        index = index.contiguous()
        self.M = int(index.shape[0])
        if num_segments is None:     # torch_scatter semantics: dim_size = index.max() + 1 (host sync)
            num_segments = int(index.max().item()) + 1 if self.M > 0 else 0
        self.S = int(num_segments)
        self.device = index.device
        self._keep = index
        lib = L.lib()
        self.buf = torch.empty(max(int(lib.yolat_segments_ints(self.M, self.S)), 1), dtype=torch.int32,
                               device=self.device)
        self._side = side_stream
        if side_stream is not None:
            side_stream.wait_stream(torch.cuda.current_stream(self.device))
            st = side_stream.cuda_stream
        else:
            st = L.stream()
        L.check(lib.yolat_segments_build(index.data_ptr(), self.M, self.S, self.buf.data_ptr(), st), 'segments_build')

    def join(self):
        """The current stream continues after the build (no-op unless built on a side stream)."""
        if self._side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
            self._side = None

    def ptr(self):
        return self.buf.data_ptr()


_graph_cache = {}


def graph_for(edge_index, num_nodes):
    """CSRGraph for `edge_index`, cached on (storage, offset, shape, strides, version) -- the second
    GraphConv layer of a step re-uses the first layer's build.  A CSRGraph passes through unchanged."""
    if isinstance(edge_index, CSRGraph):
        return edge_index
    key = (edge_index.untyped_storage().data_ptr(), edge_index.storage_offset(), tuple(edge_index.shape),
           tuple(edge_index.stride()), edge_index._version, int(num_nodes), edge_index.device)
    hit = _graph_cache.get('last')
    if hit is not None and hit[0] == key:
        return hit[1]
    g = CSRGraph(edge_index, num_nodes)
    _graph_cache['last'] = (key, g)
    return g
