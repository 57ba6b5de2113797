"""Oracle (test infrastructure): graph construction on CPU.

Restates utils/graph.py:6-124 and data/mot_graph.py:195-221,283-316 of the reference in
CPU PyTorch.  See ``oracle/__init__.py`` for who may import this.
"""
import numpy as np
import torch
import torch.nn.functional as F


def time_valid_pairs(frame_num, max_frame_dist='max'):
    """All (i, j), i < j, whose frames differ and are at most ``max_frame_dist`` apart.
    Order: ascending i, then ascending j (row-major scan of the N x N condition).
    reference: utils/graph.py:6-37
    """
    f = torch.as_tensor(frame_num).view(-1)
    gap = (f[:, None] - f[None, :]).abs()
    ok = gap > 0
    if max_frame_dist != 'max':
        ok &= gap <= int(max_frame_dist)
    ok = torch.triu(ok, diagonal=1)          # keep row < col only
    i, j = torch.nonzero(ok, as_tuple=True)  # row-major == sorted by (i, j)
    return torch.stack((i, j))


def pair_reid_dist(reid, pairs, chunk=50000):
    """||a - b + 1e-6||_2 per pair in fp32, evaluated in chunks of 50 000 pairs.
    reference: data/mot_graph.py:211 (training) and :299-303 (chunked, inference)
    """
    out = []
    for s in range(0, pairs.shape[1], chunk):
        a = reid[pairs[0, s:s + chunk]]
        b = reid[pairs[1, s:s + chunk]]
        out.append(F.pairwise_distance(a, b))
    if not out:
        return reid.new_zeros((0,))
    return torch.cat(out)


def knn_keep_mask(pwise_dist, edge_ixs, num_nodes, top_k_nns, reciprocal_k_nns=False,
                  symmetric_edges=True):
    """Which edges survive KNN pruning.  Edge (i, j) is kept iff j is among the
    ``top_k_nns`` closest neighbours of i AND (reciprocal) / OR (otherwise) i is among
    those of j; a neighbour's rank is its position in the ascending sort of the whole
    dense distance row (missing pairs = +inf).
    reference: utils/graph.py:40-87
    """
    n = int(num_nodes)
    r, c = edge_ixs[0].long(), edge_ixs[1].long()
    dense = torch.full((n, n), float('inf'), dtype=torch.float32)
    dense[r, c] = pwise_dist.view(-1).float()
    if not symmetric_edges:
        dense[c, r] = pwise_dist.view(-1).float()
    order = torch.argsort(dense, dim=1, descending=False, stable=True)   # graph.py:65 (+ stable)
    rank = torch.empty_like(order)
    rank.scatter_(1, order, torch.arange(n).expand(n, n).contiguous())  # graph.py:69-70
    near = rank < int(top_k_nns)
    near = (near & near.T) if reciprocal_k_nns else (near | near.T)  # graph.py:73-78
    return near[r, c]                                                # graph.py:85


def edge_geometry(pairs, det_cols, fps):
    """Five geometric features per pair, fp32.  ``det_cols`` maps the detection-table
    column names to float64 arrays (the DataFrame ``.values``).
    reference: utils/graph.py:90-124
    """
    i, j = pairs[0].long(), pairs[1].long()

    def col(name):
        return torch.from_numpy(np.asarray(det_cols[name])).float()

    secs = col('frame') / fps
    h, w, fx, fy = col('bb_height'), col('bb_width'), col('feet_x'), col('feet_y')
    hbar = (h[i] + h[j]) / 2
    return {
        'secs_time_dists': secs[j] - secs[i],
        'norm_feet_x_dists': (fx[j] - fx[i]) / hbar,
        'norm_feet_y_dists': (fy[j] - fy[i]) / hbar,
        'bb_height_dists': torch.log(h[j] / h[i]),
        'bb_width_dists': torch.log(w[j] / w[i]),
    }


def build_graph(frame_num, reid, det_cols, fps, dataset_params, inference_mode=False,
                max_frame_dist=None):
    """Edge construction + graph assembly for one window.

    Training mode prunes with KNN here (``symmetric_edges=False``); inference mode keeps
    every time-valid pair (the tracker prunes per sliding window later) and also returns
    the per-edge ReID distances.
    reference: data/mot_graph.py:195-221 (_get_edge_ixs), :283-316 (construct_graph_object)
    Returns dict(edge_index [2,E] int64, edge_attr [E,F] fp32, reid_emb_dists [E] or None).
    """
    mfd = dataset_params['max_frame_dist'] if max_frame_dist is None else max_frame_dist
    pairs = time_valid_pairs(frame_num, mfd)
    k = dataset_params['top_k_nns']
    if not inference_mode and k is not None:
        d = pair_reid_dist(reid, pairs, chunk=1 << 62)                # mot_graph.py:211
        keep = knn_keep_mask(d, pairs, reid.shape[0], k,
                             reciprocal_k_nns=dataset_params['reciprocal_k_nns'],
                             symmetric_edges=False)
        pairs = pairs[:, keep]
    geo = edge_geometry(pairs, det_cols, fps)
    names = [n for n in dataset_params['edge_feats_to_use'] if n in geo]
    feats = torch.stack([geo[n] for n in names]).T                    # mot_graph.py:295-296
    emb = pair_reid_dist(reid, pairs).view(-1, 1)                     # mot_graph.py:299-303
    if 'emb_dist' in dataset_params['edge_feats_to_use']:
        feats = torch.cat((feats, emb), dim=1)
    edge_attr = torch.cat((feats, feats), dim=0)                      # mot_graph.py:311
    edge_index = torch.cat((pairs, pairs.flip(0)), dim=1)             # mot_graph.py:312
    dists = torch.cat((emb, emb)).view(-1) if inference_mode else None
    return {'edge_index': edge_index, 'edge_attr': edge_attr, 'reid_emb_dists': dists}


def prune_window(edge_index, edge_attr, reid_emb_dists, num_nodes, dataset_params):
    """Per-window pruning the tracker applies before the forward pass.
    reference: tracker/mpn_tracker.py:107-112
    """
    keep = knn_keep_mask(reid_emb_dists, edge_index, num_nodes, dataset_params['top_k_nns'],
                         reciprocal_k_nns=dataset_params['reciprocal_k_nns'],
                         symmetric_edges=True)
    return edge_index[:, keep], edge_attr[keep], keep


def assign_edge_labels(edge_index, node_ids, mode='closest'):
    """Edge labels of the network-flow formulation: 1 for directed edges between detections of the same
    identity (-1 = false positive, never linked); 'closest' keeps, per source node, only the edge to the
    same-identity partner that is closest by node index among its later resp. earlier neighbours
    (scatter_min arg-min per node and direction).  reference: data/mot_graph.py:223-262"""
    ids = torch.as_tensor(node_ids).long()
    r, c = edge_index[0].long(), edge_index[1].long()
    same = (ids[r] == ids[c]) & (ids[r] != -1)
    labels = torch.zeros(edge_index.shape[1])
    if mode == 'all':
        labels[same] = 1
        return labels
    n = ids.numel()
    for future in (True, False):
        m = same & ((r < c) if future else (r > c))
        dist = (r - c).abs()
        best = {}
        for e in torch.nonzero(m).view(-1).tolist():                # first minimum wins, as scatter_min's arg-min
            k = int(r[e])
            if k not in best or int(dist[e]) < best[k][0]:
                best[k] = (int(dist[e]), e)
        for _, e in best.values():
            labels[e] = 1
    return labels
