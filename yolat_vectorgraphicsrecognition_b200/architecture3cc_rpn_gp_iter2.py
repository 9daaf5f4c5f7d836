"""Drop-in for `cad_recognition/architecture3cc_rpn_gp_iter2.py`: `Backbone`, `SparseCADGCN`,
`DetectionLoss` with the reference's constructor / forward / predict signatures, module attribute paths
and state-dict keys, executing on the sm_100a kernels behind
include/yolat_b200.h.  Reference line numbers below refer to that file.

What changes under the same surface:
  * the graph (CSR by target / by source) and the proposal segments are built once per forward and
    shared by all conv layers and by backward;
  * `fusion_block` -> `cat` -> `scatter(max)` (:62-63,122) is one fused op: out_feat [N,1152] is never built;
  * there is no host synchronisation inside forward (the reference's scatter reads `index.max()`; the
    number of proposals is taken from `data.bbox.shape[0]` instead).
"""
import numpy as np
import torch
from torch.nn import Linear as Lin

from . import ops
from .gcn_lib.sparse import MultiSeq, MLP, GraphConv, ResBlock
from .graph import CSRGraph, Segments
from .torch_scatter import scatter


def _packed_to_device(data, device):
    """A batch.PackedBatch reaches the device as ONE copy of its pinned buffer; per-image offsets deferred by
    `batch.collate(...).defer_offsets()` (train.py:238-258) are applied there.  The device views are remembered on the
    batch so that DetectionLoss finds the labels without a second upload."""
    from .batch import apply_offsets
    buf, ns = data.device_twin(device)
    buf.copy_(data.host, non_blocking=True)
    if getattr(data, 'offsets_pending', False):
        apply_offsets(ns)
    data._device = ns
    return ns


def _dev(t, device):
    """`.cuda()` of the reference (:107-115), a no-op for tensors that already live on the device."""
    return t if t.is_cuda else t.to(device, non_blocking=True)


_AUX_STREAMS = {}


def _aux_stream(dev):
    """One library-owned side stream per device (module state, not model state: models stay deep-copyable)."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _AUX_STREAMS:
        _AUX_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _AUX_STREAMS[key]


class Backbone(torch.nn.Module):
    """:15-71"""

    def __init__(self, opt, n_edges=3, edge_max_pool=torch.nn.AdaptiveAvgPool1d):
        super(Backbone, self).__init__()
        channels = opt.n_filters
        act = opt.act
        norm = opt.norm
        bias = opt.bias
        conv = 'attr_edge_gp2'  # opt.conv is ignored by the reference (:22)
        c_growth = channels
        n_edges = 1
        self.n_edges = n_edges

        self.n_blocks = opt.n_blocks
        self.n_blocks_out = opt.n_blocks_out
        self.heads = torch.nn.ModuleList()
        self.n_classes = opt.n_classes
        self.class_specific = opt.class_specific

        self.head = GraphConv(opt.in_channels, channels, conv, act, norm, bias)
        self.backbone = MultiSeq(*[ResBlock(channels, conv, act, norm, bias) for i in range(self.n_blocks - 1)])
        fusion_dims = int(channels + c_growth * (self.n_blocks_out - 1))
        self.fusion_block = MLP([fusion_dims, 1024], act, norm, bias)
        self.fusion_block_super = MLP([fusion_dims, 1024], act, norm, bias)
        self.fusion_dims = fusion_dims

    # -- shared trunk: the conv stack (:45-57) and the feature-map selection (:60-61,65-66) -----------------
    def _trunk(self, x, edges, edge_weights, edge_attrs):
        graph = edges[0] if isinstance(edges[0], CSRGraph) else CSRGraph(edges[0], x.shape[0])
        f, f_super = self.head(x, graph, edge_weights[0], edge_attrs[0], x_node=x)
        feats, feats_super = [f], [f_super]
        for i in range(self.n_blocks - 1):
            f, f_super = self.backbone[i](feats[-1], graph, edge_weights[0], edge_attrs[0], x_node=feats_super[-1])
            feats.append(f)
            feats_super.append(f_super)
        sel = range(self.n_blocks - self.n_blocks_out, self.n_blocks)
        feats = torch.cat([feats[i] for i in sel], dim=1)
        feats_super = torch.cat([feats_super[i] for i in sel], dim=1)
        return feats, feats_super

    def _super_branch(self, feats_super, seg):
        feats_super = scatter(feats_super, seg, dim=0, reduce='mean')                   # :67
        fusion_feats_super = self.fusion_block_super(feats_super)                        # :68
        return torch.cat((fusion_feats_super, feats_super), dim=1)                       # :69

    def forward(self, x, edges, edge_weights, edge_attrs, bbox_idx):
        """Reference signature (:44); materialises out_feat [N, 1024 + fusion_dims] like the reference."""
        feats, feats_super = self._trunk(x, edges, edge_weights, edge_attrs)
        fusion_feats = self.fusion_block(feats)                                           # :62
        out_feat = torch.cat((fusion_feats, feats), dim=1)                                # :63
        seg = bbox_idx if isinstance(bbox_idx, Segments) else Segments(bbox_idx)
        return out_feat, self._super_branch(feats_super, seg)

    def forward_pooled(self, x, edges, edge_weights, edge_attrs, seg):
        """Fused variant used by SparseCADGCN.forward: returns scatter(out_feat, 'max') [B, 1024 + fusion_dims]
        (:62-63 + :122 in one op) and out_feat_super."""
        feats, feats_super = self._trunk(x, edges, edge_weights, edge_attrs)
        seg.join()                      # built on a side stream while the GraphConv stack ran (SparseCADGCN.forward)
        stages = self.fusion_block.stages()
        if len(stages) == 1 and stages[0][0] == 'stage' and stages[0][2] is not None and stages[0][3]:
            _, lin, bn, _ = stages[0]
            pooled = ops.FuseMaxFn.apply(seg, bn.training, (bn.running_mean, bn.running_var, bn.num_batches_tracked),
                                         feats, lin.weight, lin.bias, bn.weight, bn.bias)
        else:   # non-default act / norm: unfused composition
            pooled = scatter(torch.cat((self.fusion_block(feats), feats), dim=1), seg, dim=0, reduce='max')
        return pooled, self._super_branch(feats_super, seg)


class SparseCADGCN(torch.nn.Module):
    """:73-356"""

    def __init__(self, opt, n_edges=3, edge_max_pool=torch.nn.AdaptiveAvgPool1d, expand_ratio=0.25):
        super(SparseCADGCN, self).__init__()
        self.expand_ratio = expand_ratio
        act = opt.act
        norm = opt.norm
        bias = opt.bias
        self.n_classes = opt.n_classes
        self.classifier = opt.classifier
        self.class_specific = opt.class_specific
        self.dim_stat = 0

        import os
        self.check_graph = os.environ.get('YOLAT_CHECK_GRAPH', '0') not in ('', '0')
        self._in_predict = False
        self.cls_net = Backbone(opt)
        self.prediction_cls = MultiSeq(*[
            MLP([(self.cls_net.fusion_dims + 1024) * 2 + self.dim_stat, 512], act, norm, bias),
            MLP([512, 256], act, norm, bias, drop=opt.dropout),
            MLP([256, opt.n_classes], None, None, bias)])
        self.model_init()

    def model_init(self):
        """:97-104 -- kaiming_normal_ on every Linear weight, zero biases (same module order => same RNG stream)."""
        for m in self.modules():
            if isinstance(m, Lin):
                torch.nn.init.kaiming_normal_(m.weight)
                m.weight.requires_grad = True
                if m.bias is not None:
                    m.bias.data.zero_()
                    m.bias.requires_grad = True

    def _device(self):
        return next(self.parameters()).device

    def forward(self, data, slices):
        """:106-137.  `data` carries x [N,Cin], bbox_idx [N], edge [E,2], bbox [B,4], stat_feats, e_attr [E,4]
        as CPU or CUDA tensors; returns (pred_cls [B,ncls], pred_bbox [B,4])."""
        dev = self._device()
        if hasattr(data, 'device_twin') and hasattr(data, 'host'):      # batch.PackedBatch
            data = _packed_to_device(data, dev)
        x = _dev(data.x, dev)
        bbox_idx = _dev(data.bbox_idx, dev)
        edge = _dev(data.edge, dev)
        pred_bbox = _dev(data.bbox, dev)
        e_attr = _dev(data.e_attr, dev)
        # stat_feats is copied by the reference (:112) but unused (dim_stat = 0, :87): not moved here.

        # edges = [data.edge.cuda().T]  (:110).  Edge endpoints outside [0, N) make the reference fail (index error in PyG's
        # gather, KeyError in build_data); the graph build counts them on the device, and the count is read back -- one
        # host sync -- wherever a sync exists anyway (predict) or when `check_graph` is set (YOLAT_CHECK_GRAPH=1).
        graph = CSRGraph(edge.T, x.shape[0], check=self.check_graph or self._in_predict)
        # one proposal per bbox row: no index.max() sync; built beside the GraphConv stack, joined in forward_pooled
        seg = Segments(bbox_idx, pred_bbox.shape[0], side_stream=_aux_stream(dev))
        pooled, out_feat_cls_super = self.cls_net.forward_pooled(x, [graph], [None], [e_attr], seg)   # :121-122
        out_feat_cls = torch.cat([pooled, out_feat_cls_super], dim=1)                                  # :127
        pred_cls = self.prediction_cls(out_feat_cls)                                                   # :128
        if self.classifier != 'softmax':
            pred_cls = torch.sigmoid(pred_cls)                                                         # :132-133
        return pred_cls, pred_bbox

    # ----------------------------------------------------------------------------------------------
    # two-stage inference (:139-356): classify the root proposal of every connected component, expand
    # the children of the roots classified as background, classify those, interleave.
    # ----------------------------------------------------------------------------------------------
    def _device_data(self, data):
        """The batch on the model's device, moved once per predict() (the reference re-slices on the host and copies
        every slice: :167-234, :107-115)."""
        from types import SimpleNamespace
        dev = self._device()
        out = SimpleNamespace()
        for k in ('x', 'pos', 'bbox_idx', 'edge', 'e_attr', 'bbox', 'stat_feats'):
            v = getattr(data, k, None)
            setattr(out, k, None if v is None else _dev(v, dev).contiguous())
        return out

    @staticmethod
    def _build_data_device(dd, nodes, slices):
        """The range lists + `build_data` of the reference (:153-234) on the device (csrc/slicing.cu): the K (start, length) pairs of the selected
        proposals are the only host data; range expansion, the old->new renumbering of edge endpoints and the dense
        bbox_idx renumbering are kernels, the row gathers are index_selects.  Returns (batch namespace, slice_bbox)."""
        from types import SimpleNamespace
        from . import _lib as L
        lib = L.lib()
        dev = dd.x.device
        st = torch.cuda.current_stream(dev).cuda_stream
        K = len(nodes)
        tab = np.zeros((4, K + 1), dtype=np.int64)          # pos start | pos prefix | edge start | edge prefix
        bbox = []
        for k, (node, i) in enumerate(nodes):
            v = node.value
            tab[0, k] = v['idx_pos'][0] + int(slices['pos'][i])
            tab[1, k + 1] = tab[1, k] + (v['idx_pos'][1] - v['idx_pos'][0])
            tab[2, k] = v['idx_edge'][0] + int(slices['edge'][i])
            tab[3, k + 1] = tab[3, k] + (v['idx_edge'][1] - v['idx_edge'][0])
            bbox.append(int(v['idx_bbox'] + slices['bbox'][i]))
        Np, Ep = int(tab[1, K]), int(tab[3, K])
        t = torch.from_numpy(tab).to(dev)
        sp = torch.empty(Np, dtype=torch.long, device=dev)
        se = torch.empty(Ep, dtype=torch.long, device=dev)
        L.check(lib.yolat_expand_ranges(t[0].data_ptr(), t[1].data_ptr(), K, Np, L.ptr(sp), st), 'yolat_expand_ranges')
        L.check(lib.yolat_expand_ranges(t[2].data_ptr(), t[3].data_ptr(), K, Ep, L.ptr(se), st), 'yolat_expand_ranges')
        N_all, E_all = dd.x.shape[0], dd.edge.shape[0]
        ws = torch.empty(int(lib.yolat_slice_graph_ints(N_all, Np)), dtype=torch.int32, device=dev)
        nd = SimpleNamespace()
        nd.edge = torch.empty(Ep, 2, dtype=torch.long, device=dev)
        nd.bbox_idx = torch.empty(Np, dtype=torch.long, device=dev)
        L.check(lib.yolat_slice_graph(L.ptr(sp), Np, L.ptr(se), Ep, L.ptr(dd.edge), E_all, N_all, L.ptr(dd.bbox_idx),
                                      L.ptr(ws), L.ptr(nd.edge), L.ptr(nd.bbox_idx), st), 'yolat_slice_graph')
        sb = torch.as_tensor(bbox, dtype=torch.long, device=dev)
        nd.x = dd.x.index_select(0, sp)
        nd.pos = dd.pos.index_select(0, sp) if dd.pos is not None else None
        nd.e_attr = dd.e_attr.index_select(0, se)
        nd.bbox = dd.bbox.index_select(0, sb)
        nd.stat_feats = dd.stat_feats.index_select(0, sb) if dd.stat_feats is not None else None
        return nd, bbox

    def predict(self, data, slices):
        roots = data.roots
        slice_root = slices['roots']
        root_nodes, slice_image_bbox_root = [], [0]
        for i in range(0, len(slice_root) - 1):
            for root in roots[slice_root[i]:slice_root[i + 1]]:
                root_nodes.append((root, i))
            slice_image_bbox_root.append(len(root_nodes))
        dd = self._device_data(data)
        nd, slice_bbox = self._build_data_device(dd, root_nodes, slices)
        self._in_predict = True          # forward() verifies the edge endpoints: predict syncs with the host anyway
        try:
            return self._predict_stages(dd, nd, slice_bbox, slices, roots, slice_root, slice_image_bbox_root)
        finally:
            self._in_predict = False

    def _predict_stages(self, dd, nd, slice_bbox, slices, roots, slice_root, slice_image_bbox_root):
        pred_cls, pred_bbox = self.forward(nd, slices)

        _, is_object = pred_cls.max(1)
        has_object = (is_object == self.n_classes - 1).cpu()
        slice_bbox_root = slice_bbox

        child_nodes, slice_image_bbox_child = [], [0]
        count = 0
        for i in range(0, len(slice_root) - 1):
            for root in roots[slice_root[i]:slice_root[i + 1]]:
                if has_object[count]:
                    for child in root.children:
                        child_nodes.append((child, i))
                count += 1
            slice_image_bbox_child.append(len(child_nodes))
        n_child_pos = sum(c.value['idx_pos'][1] - c.value['idx_pos'][0] for c, _ in child_nodes)

        if n_child_pos == 0:                                   # `len(slice_pos) == 0` (:299)
            slice_image_bbox = slice_image_bbox_root
            slice_bbox = slice_bbox_root
        else:
            nd2, slice_bbox = self._build_data_device(dd, child_nodes, slices)
            pred_cls2, pred_bbox2 = self.forward(nd2, slices)

            def interleaf_pc(slice_p, slice_c, out_p, out_c):
                out, s = [], [0]
                for i in range(len(slice_c) - 1):
                    out.append(out_p[slice_p[i]:slice_p[i + 1]])
                    out.append(out_c[slice_c[i]:slice_c[i + 1]])
                    s.append(s[-1] + slice_p[i + 1] - slice_p[i] + slice_c[i + 1] - slice_c[i])
                return out, s

            pred_cls, slice_image_bbox = interleaf_pc(slice_image_bbox_root, slice_image_bbox_child, pred_cls, pred_cls2)
            pred_bbox, slice_image_bbox = interleaf_pc(slice_image_bbox_root, slice_image_bbox_child, pred_bbox,
                                                       pred_bbox2)
            slice_bbox, slice_image_bbox = interleaf_pc(slice_image_bbox_root, slice_image_bbox_child,
                                                        torch.tensor(slice_bbox_root), torch.tensor(slice_bbox))
            pred_cls = torch.cat(pred_cls, dim=0)
            pred_bbox = torch.cat(pred_bbox, dim=0)
            slice_bbox = torch.cat(slice_bbox, dim=0)

        w = (pred_bbox[:, 2] - pred_bbox[:, 0]) * 1.05
        h = (pred_bbox[:, 3] - pred_bbox[:, 1]) * 1.05
        center_x = (pred_bbox[:, 2] + pred_bbox[:, 0]) / 2
        center_y = (pred_bbox[:, 3] + pred_bbox[:, 1]) / 2
        pred_bbox = torch.stack([center_x - w / 2, center_y - h / 2, center_x + w / 2, center_y + h / 2], dim=1)
        return pred_cls, pred_bbox, None, slice_bbox, slice_image_bbox, None


class DetectionLoss(torch.nn.Module):
    """:358-379"""

    def __init__(self, opt):
        super(DetectionLoss, self).__init__()
        self.classifier = opt.classifier
        self.cls_loss = None if opt.classifier == 'softmax' else torch.nn.BCELoss()

    def forward(self, out, data):
        pred_cls = out[0]
        gt_cls = _dev(getattr(data, '_device', data).labels, pred_cls.device)
        if self.classifier == 'softmax':
            l0 = ops.softmax_cross_entropy(pred_cls, gt_cls)      # CrossEntropyLoss(mean), :363,376
        else:
            gt = torch.zeros(pred_cls.size(), device=pred_cls.device).scatter_(1, gt_cls.unsqueeze(1), 1)
            l0 = self.cls_loss(pred_cls, gt)
        return {'loss': l0, 'loss_cls': l0}
