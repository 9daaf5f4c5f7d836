"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's proposal enumeration.

Follows `Datasets/graph_dict3.py:309-789` (`SESYDFloorPlan._get_proposal`, do_mixup off) step by step in plain python /
numpy -- literal `move_endpoint` while-loops, `np.arange` grids, python sets for the node sets, `np.mean` / `np.std`
for the statistics -- and `utils/det_util.py:311-362` for the IoU / IoS / overlap arithmetic.  The one deliberate
difference from the reference: the de-duplicated node sets of a component are kept in first-occurrence order of the
window walk, where the reference iterates a python `set` of tuples (`:557`, CPython hash-table order).  Pinned
against the UNMODIFIED reference by `oracle/make_golden_proposals.py` -> `tests/golden/proposals_*.pkl` (compared
after `canonical_order`, which sorts each component's proposals by their box -- unique within a component).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np


class idxTree(object):
    def __init__(self):
        self.children = []
        self.value = {}


def _move_endpoint(x, values, bound):          # graph_dict3.py:482-490
    if x >= len(values):
        return x - 1
    while values[x] <= bound:
        x += 1
        if x >= len(values):
            break
    return x - 1


def _move_endpoint_close(x, values, bound):    # graph_dict3.py:492-500
    if x >= len(values):
        return x - 1
    while values[x] < bound:
        x += 1
        if x >= len(values):
            break
    return x - 1


def _windows(x_values, y_values, x_grids, y_grids):
    """Rank windows (x0, y0, x1, y1) in the order the four nested loops of :502-523 reach them."""
    out = []
    prev_y0 = -1
    for iy0, gy0 in enumerate(y_grids):
        y0 = _move_endpoint_close(prev_y0 + 1, y_values, gy0)
        if y0 != len(y_values):
            y0 += 1
        if y0 == prev_y0:
            continue
        prev_y0 = y0
        prev_x0 = -1
        for ix0, gx0 in enumerate(x_grids):
            x0 = _move_endpoint_close(prev_x0 + 1, x_values, gx0)
            if x0 != len(y_values):            # sic (:516)
                x0 += 1
            if x0 == prev_x0:
                continue
            prev_x0 = x0
            prev_y1 = y0
            for gy1 in y_grids[iy0 + 1:]:
                y1 = _move_endpoint(prev_y1 + 1, y_values, gy1)
                if y1 == prev_y1:
                    continue
                prev_y1 = y1
                prev_x1 = x0
                for gx1 in x_grids[ix0 + 1:]:
                    x1 = _move_endpoint(prev_x1 + 1, x_values, gx1)
                    if x1 == prev_x1:
                        continue
                    prev_x1 = x1
                    out.append((x0, y0, x1, y1))
    return out


def _iou_ios(box, gts):                        # utils/det_util.py:326-341 (box1 = one proposal, box2 = gts)
    ix1 = np.maximum(box[0], gts[:, 0])
    iy1 = np.maximum(box[1], gts[:, 1])
    ix2 = np.minimum(box[2], gts[:, 2])
    iy2 = np.minimum(box[3], gts[:, 3])
    inter = np.maximum(ix2 - ix1, 0) * np.maximum(iy2 - iy1, 0)
    a1 = (box[2] - box[0]) * (box[3] - box[1])
    a2 = (gts[:, 2] - gts[:, 0]) * (gts[:, 3] - gts[:, 1])
    return inter / (a1 + a2 - inter + 1e-16), inter / a2


def get_proposal(graph_dict, gt_bbox, gt_labels, bbox_sampling_step=5, n_classes=17, normalize_bbox=True):
    cc = graph_dict['cc']
    pos = np.asarray(graph_dict['pos']['spatial'], dtype=np.float64)
    edge = np.asarray(graph_dict['edge']['shape'])
    edge_super = np.asarray(graph_dict['edge']['super'])
    e_attr = np.asarray(graph_dict['edge_attr']['shape'])
    e_attr_super = np.asarray(graph_dict['edge_attr']['super'])
    is_super = np.asarray(graph_dict['attr']['is_super'])
    is_control = np.asarray(graph_dict['attr']['is_control'])
    gt_bbox = np.asarray(gt_bbox, dtype=np.float64)

    # (1) control points out, everything renumbered (:325-352)
    keep = np.flatnonzero(is_control.reshape(-1) == 0)
    o2n = {int(o): n for n, o in enumerate(keep)}
    edge = np.array([[o2n[int(a)], o2n[int(b)]] for a, b in edge], dtype=np.int64).reshape(-1, 2)
    edge_super = np.array([[o2n[int(a)], o2n[int(b)]] for a, b in edge_super], dtype=np.int64).reshape(-1, 2)
    cc = [[o2n[int(i)] for i in cluster] for cluster in cc]
    pos = pos[keep]
    is_super = is_super[keep]

    def pair_lists(edges):                     # the dense adjacency of :560-571, as a dict of unordered pairs
        adj = {}
        for i, (a, b) in enumerate(edges.tolist()):
            adj.setdefault((a, b), []).append(i)
            if a != b:
                adj.setdefault((b, a), []).append(i)
            else:                              # A[a][a] gets the edge twice; never read (pairs are i < j)
                adj[(a, a)].append(i)
        return adj

    adj_shape, adj_super = pair_lists(edge), pair_lists(edge_super)

    out = {k: [] for k in ('pos', 'is_super', 'edge', 'edge_super', 'e_attr', 'e_attr_super', 'labels', 'bbox',
                           'targets', 'bbox_idx', 'stats', 'has_obj')}
    sl_pos, sl_edge, sl_super, sl_bbox = [0], [0], [0], [0]
    roots, offset, bbox_count = [], 0, 0

    for cluster in cc:
        pc = pos[cluster, :]
        max_x, min_x, max_y, min_y = pc[:, 0].max(), pc[:, 0].min(), pc[:, 1].max(), pc[:, 1].min()
        x_values = sorted(set(pc[:, 0].tolist()))
        y_values = sorted(set(pc[:, 1].tolist()))
        xr = {v: i for i, v in enumerate(x_values)}
        yr = {v: i for i, v in enumerate(y_values)}
        rx = [xr[v] for v in pc[:, 0].tolist()]
        ry = [yr[v] for v in pc[:, 1].tolist()]
        x_step = (max_x - min_x) / bbox_sampling_step
        y_step = (max_y - min_y) / bbox_sampling_step
        x_grids = np.append(np.arange(min_x, max_x, x_step), max_x)      # ValueError('arange: cannot compute length') on a zero step
        y_grids = np.append(np.arange(min_y, max_y, y_step), max_y)

        # (2) node sets of the windows; d00[y1][x1] - d00[y1][x0-1] - d00[y0-1][x1] (:544-551) = ranks inside the window
        seen, sub_clusters = set(), []
        for x0, y0, x1, y1 in _windows(x_values, y_values, x_grids.tolist(), y_grids.tolist()):
            members = tuple(sorted(cluster[i] for i in range(len(cluster))
                                   if x0 <= rx[i] <= x1 and y0 <= ry[i] <= y1))
            if members not in seen:
                seen.add(members)
                sub_clusters.append(members)

        cc_box = np.array([min_x, min_y, max_x, max_y])
        ix1, iy1 = np.maximum(cc_box[0], gt_bbox[:, 0]), np.maximum(cc_box[1], gt_bbox[:, 1])
        ix2, iy2 = np.minimum(cc_box[2], gt_bbox[:, 2]), np.minimum(cc_box[3], gt_bbox[:, 3])
        valid = np.where((ix2 > ix1) & (iy2 > iy1))[0]                     # det_util.py:355-362
        if valid.shape[0] == 0:
            print('cc has no intersect gt bbox')
            raise SystemExit

        first_of_cc = bbox_count
        for idxs in sub_clusters:
            local = {v: i for i, v in enumerate(idxs)}
            e_ids, s_ids = [], []
            for i in range(len(idxs)):
                for j in range(i + 1, len(idxs)):
                    e_ids += adj_shape.get((idxs[i], idxs[j]), [])
                    s_ids += adj_super.get((idxs[i], idxs[j]), [])
            if len(e_ids) == 0:
                continue
            pb = pos[list(idxs), :]
            e_loc = np.array([[local[a], local[b]] for a, b in edge[e_ids].tolist()], dtype=np.int64)
            s_loc = np.array([[local[a], local[b]] for a, b in edge_super[s_ids].tolist()], dtype=np.int64).reshape(-1, 2)
            bx1, bx0, by1, by0 = pb[:, 0].max(), pb[:, 0].min(), pb[:, 1].max(), pb[:, 1].min()
            if bx1 - bx0 < 1e-4 or by1 - by0 < 1e-4:
                continue
            box = np.array([bx0, by0, bx1, by1])
            iou, ios = _iou_ios(box, gt_bbox[valid, :])
            g = int(np.argmax(iou))
            if iou[g] > 0.7:
                label, target = gt_labels[valid[g]], gt_bbox[valid[g]]
            else:
                label, target = n_classes - 1, np.zeros(4)
            has_obj = 1 if ios[g] > 0.7 else 0

            nbrs = [set() for _ in idxs]
            for a, b in e_loc.tolist():
                nbrs[a].add(b)
                nbrs[b].add(a)
            dots, less, right, more = [], 0, 0, 0
            for anchor, ns in enumerate(nbrs):
                ns = sorted(ns)
                for i in range(len(ns)):
                    for j in range(i + 1, len(ns)):
                        v0, v1 = pb[ns[i]] - pb[anchor], pb[ns[j]] - pb[anchor]
                        dot = v0[0] * v1[0] + v0[1] * v1[1]
                        if dot <= -1e-2:
                            more += 1
                        elif dot >= 1e-2:
                            less += 1
                        elif abs(dot) < 1e-2:
                            right += 1
                        dots.append(dot)
            if len(dots) == 0:
                continue
            dots = np.array(dots)
            ea = e_attr[e_ids]
            stats = [len(idxs), len(e_ids), right, less, more, bx1 - bx0, by1 - by0, np.mean(dots), np.max(dots),
                     np.min(dots), np.std(dots), np.mean(ea[:, -1]), np.std(ea[:, -1])]
            if normalize_bbox:
                pb = (pb - [bx0, by0]) / [bx1 - bx0, by1 - by0]

            sl_pos.append(sl_pos[-1] + len(idxs))
            sl_edge.append(sl_edge[-1] + len(e_ids))
            sl_super.append(sl_super[-1] + len(s_ids))
            sl_bbox.append(sl_bbox[-1] + 1)
            out['pos'].append(pb)
            out['is_super'].append(is_super[list(idxs)].reshape(len(idxs), -1))
            out['edge'].append(e_loc + offset)
            if len(s_ids):
                out['edge_super'].append(s_loc + offset)
            out['e_attr'].append(ea)
            out['e_attr_super'].append(e_attr_super[s_ids])
            out['labels'].append(label)
            out['has_obj'].append(has_obj)
            out['bbox_idx'] += [bbox_count] * len(idxs)
            out['bbox'].append([bx0, by0, bx1, by1])
            out['targets'].append(np.asarray(target, dtype=np.float64).reshape(1, 4))
            out['stats'].append(np.array(stats, dtype=np.float64).reshape(1, -1))
            offset += len(idxs)
            bbox_count += 1

        # (4) the largest box of the component is the root (:724-750)
        boxes = np.array(out['bbox'])[first_of_cc:]
        area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
        root_i = first_of_cc + int(np.argmax(area))

        def node(i):
            t = idxTree()
            t.value['idx_pos'] = (sl_pos[i], sl_pos[i + 1])
            t.value['idx_edge'] = (sl_edge[i], sl_edge[i + 1])
            t.value['idx_edge_super'] = (sl_super[i], sl_super[i + 1])
            t.value['idx_bbox'] = sl_bbox[i]
            return t

        root = node(root_i)
        root.children = [node(i) for i in range(first_of_cc, bbox_count) if i != root_i]
        roots.append(root)

    pos_o = np.concatenate(out['pos'], axis=0)
    return (pos_o, np.concatenate(out['is_super'], axis=0), np.zeros((pos_o.shape[0], 1)),
            np.concatenate(out['edge'], axis=0), np.concatenate(out['edge_super'], axis=0),
            np.concatenate(out['e_attr'], axis=0), np.concatenate(out['e_attr_super'], axis=0), out['labels'],
            np.array(out['bbox_idx']), np.array(out['bbox']), np.concatenate(out['targets'], axis=0),
            np.concatenate(out['stats'], axis=0), out['has_obj'], roots)


def canonical_order(result):
    """Reorder the proposals of each component by their box (min_x, min_y, max_x, max_y) -- unique per component --
    so that two enumerations whose per-component order differs (the reference's is a CPython hash-table artefact)
    can be compared field by field.  Returns a dict of arrays; node / edge blocks move with their proposal and their
    indices are renumbered."""
    (pos, is_super, is_control, edge, edge_super, e_attr, e_attr_super, labels, bbox_idx, bbox, targets, stats, has_obj,
     roots) = result
    bbox = np.asarray(bbox)
    b = len(labels)
    order, root_of, values = [], [], {}
    for root in roots:
        members = [root] + list(root.children)
        for m in members:
            values[m.value['idx_bbox']] = m.value
        ids = sorted((m.value['idx_bbox'] for m in members), key=lambda i: tuple(bbox[i].tolist()))
        assert len({tuple(bbox[i].tolist()) for i in ids}) == len(ids), 'two proposals of one component share a box'
        root_of.append(len(order) + ids.index(root.value['idx_bbox']))
        order += ids
    assert sorted(order) == list(range(b)), 'roots do not cover every proposal exactly once'
    new = {k: [] for k in ('pos', 'is_super', 'edge', 'edge_super', 'e_attr', 'e_attr_super', 'bbox_idx')}
    off = 0
    for new_i, old_i in enumerate(order):
        v = values[old_i]
        (p0, p1), (e0, e1), (s0, s1) = v['idx_pos'], v['idx_edge'], v['idx_edge_super']
        assert (np.asarray(bbox_idx[p0:p1]) == old_i).all()
        new['pos'].append(np.asarray(pos[p0:p1]))
        new['is_super'].append(np.asarray(is_super[p0:p1]).reshape(p1 - p0, -1))
        new['edge'].append(np.asarray(edge[e0:e1]).reshape(-1, 2) - p0 + off)
        new['edge_super'].append(np.asarray(edge_super[s0:s1]).reshape(-1, 2) - p0 + off)
        new['e_attr'].append(np.asarray(e_attr[e0:e1]))
        new['e_attr_super'].append(np.asarray(e_attr_super[s0:s1]))
        new['bbox_idx'].append(np.full((p1 - p0, 1), new_i, dtype=np.int64))
        off += p1 - p0
    res = {k: np.concatenate(v, axis=0) for k, v in new.items()}
    res['bbox_idx'] = res['bbox_idx'].reshape(-1)
    res.update({
        'labels': np.asarray(labels, dtype=np.int64)[order], 'has_obj': np.asarray(has_obj, dtype=np.int64)[order],
        'bbox': bbox[order], 'bbox_targets': np.asarray(targets)[order], 'stat_feats': np.asarray(stats)[order],
        'roots': np.asarray(root_of, dtype=np.int64), 'is_control': np.asarray(is_control),
        'per_cc': np.asarray([1 + len(r.children) for r in roots], dtype=np.int64),
    })
    return res


# ---- synthetic graph_dicts shaped like the Dataset's pickles (svg_parser.py output), for goldens and parity tests ----
def synth_graph_dict(seed, n_cc=6, max_nodes=14, grid=6, with_control=True, n_gt_extra=3, parallel_edges=True,
                     jitter=0.0, dup_points=0.0):
    """Components on disjoint patches of the unit square; node coordinates on a coarse lattice (repeated x / y values,
    as floor plans have), a spanning path plus random extra / parallel / self-loop edges per component, super edges,
    interleaved control points, 6-column edge attributes.  Ground truth: the box of every other component (IoU 1 ->
    labelled), a shrunk box (IoS high, IoU low) and a few random boxes.  `jitter` > 0 moves every node off the lattice
    by a random fraction of a cell (arbitrary doubles, mostly distinct coordinate values, as parsed SVG has);
    `dup_points` is the probability that a node sits exactly on an earlier node of its component.  Both default to off
    and draw no random numbers then, so the committed goldens' inputs do not change."""
    rng = np.random.RandomState(seed)
    pos, is_control, is_super, cc, edges, supers = [], [], [], [], [], []
    gts, gt_labels = [], []
    cols = int(np.ceil(np.sqrt(n_cc)))
    for c in range(n_cc):
        ox, oy = (c % cols) / cols, (c // cols) / cols
        k = int(rng.randint(3, max_nodes + 1))
        # distinct lattice points with at least two distinct x and two distinct y values, first three not collinear
        while True:
            cells = rng.choice(grid * grid, size=min(k, grid * grid), replace=False)
            px, py = cells % grid, cells // grid
            if len(set(px.tolist())) > 1 and len(set(py.tolist())) > 1:
                break
        ids = []
        for x, y in zip(px.tolist(), py.tolist()):
            if with_control and rng.rand() < 0.3:              # a control point in between (dropped by the o2n map)
                pos.append([rng.rand(), rng.rand()]); is_control.append(1); is_super.append(0)
            p = [ox + (0.05 + 0.9 * x / (grid - 1)) / cols, oy + (0.05 + 0.9 * y / (grid - 1)) / cols]
            if jitter > 0:
                p = [p[0] + jitter * (rng.rand() - 0.5) * 0.9 / (grid - 1) / cols,
                     p[1] + jitter * (rng.rand() - 0.5) * 0.9 / (grid - 1) / cols]
            if dup_points > 0 and len(ids) >= 3 and rng.rand() < dup_points:
                p = list(pos[ids[int(rng.randint(0, len(ids)))]])
            ids.append(len(pos))
            pos.append(p)
            is_control.append(0); is_super.append(int(rng.rand() < 0.1))
        order = rng.permutation(len(ids)).tolist()
        cc.append([ids[i] for i in order])
        for a, b in zip(ids[:-1], ids[1:]):                    # path: every inner node has two neighbours -> angles
            edges.append([a, b] if rng.rand() < 0.5 else [b, a])
        for _ in range(int(rng.randint(0, len(ids) + 1))):
            a, b = rng.choice(ids, 2)
            edges.append([int(a), int(b)])                     # may repeat an edge or be a self loop
        if parallel_edges and len(ids) > 2:
            edges.append([ids[1], ids[0]])
        for _ in range(int(rng.randint(1, 4))):
            a, b = rng.choice(ids, 2, replace=False)
            supers.append([int(a), int(b)])
        p = np.array([pos[i] for i in ids])
        box = [p[:, 0].min(), p[:, 1].min(), p[:, 0].max(), p[:, 1].max()]
        if c % 2 == 0:
            gts.append(box); gt_labels.append(int(rng.randint(0, 16)))
        else:                                                  # strictly inside: touches the component, IoS 1, IoU < 0.7
            w, h = box[2] - box[0], box[3] - box[1]
            gts.append([box[0] + 0.3 * w, box[1] + 0.3 * h, box[2] - 0.3 * w, box[3] - 0.3 * h])
            gt_labels.append(int(rng.randint(0, 16)))
    for _ in range(n_gt_extra):
        a = rng.rand(2) * 0.7
        gts.append([a[0], a[1], a[0] + 0.05 + 0.25 * rng.rand(), a[1] + 0.05 + 0.25 * rng.rand()])
        gt_labels.append(int(rng.randint(0, 16)))
    perm = rng.permutation(len(edges))
    edges = np.array(edges, dtype=np.int64)[perm]
    supers = np.array(supers, dtype=np.int64)
    graph_dict = {
        'cc': cc,
        'pos': {'spatial': np.array(pos, dtype=np.float64)},
        'edge': {'shape': edges, 'super': supers},
        'edge_attr': {'shape': rng.randn(len(edges), 6), 'super': np.zeros((len(supers), 6))},
        'attr': {'is_super': np.array(is_super, dtype=np.int64).reshape(-1, 1),
                 'is_control': np.array(is_control, dtype=np.int64).reshape(-1, 1)},
        'img_width': 1000.0, 'img_height': 1000.0,
    }
    return graph_dict, np.array(gts, dtype=np.float64), np.array(gt_labels, dtype=np.int64)
