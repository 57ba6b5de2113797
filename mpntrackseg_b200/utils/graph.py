"""Drop-in graph utilities with the reference's names and signatures, executed by the CUDA
library.  reference: src/mot_neural_solver/utils/graph.py:6-124.

The reference moves its inputs to the GPU itself when ``use_cuda`` is set; these functions
always compute on the GPU (there is no CPU implementation here) and only use ``use_cuda`` to
decide where the RESULT lives, as the reference does.
"""
import numpy as np
import torch

from .. import ops


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError('mpntrackseg_b200 needs a CUDA device (no CPU fallback)')
    return torch.device('cuda')


def get_time_valid_conn_ixs(frame_num, max_frame_dist, use_cuda, return_undirected=True):
    """Pairs of nodes in different frames at most ``max_frame_dist`` frames apart ('max' = no
    bound).  Returns a CPU LongTensor [2, E_c] with row < col, sorted by (row, col); with
    ``return_undirected=False`` the reference returns BOTH orientations as a (row, col) tuple
    on the compute device.  reference: utils/graph.py:6-37"""
    assert isinstance(max_frame_dist, (int, np.integer)) or max_frame_dist == 'max'
    f = torch.as_tensor(frame_num).to(_dev(), torch.int64).view(-1)
    pairs = ops.time_valid_pairs(f, -1 if max_frame_dist == 'max' else int(max_frame_dist))
    if not return_undirected:
        # all ordered pairs in row-major order of the N x N condition matrix
        both = torch.cat((pairs, pairs.flip(0)), dim=1)
        order = torch.argsort(both[0] * f.numel() + both[1])
        both = both[:, order]
        return both[0], both[1]
    return pairs.cpu()


def get_knn_mask(pwise_dist, edge_ixs, num_nodes, top_k_nns, use_cuda, reciprocal_k_nns=False,
                 symmetric_edges=True):
    """Bool mask [E'] of the edges that survive top-k (reciprocal) nearest-neighbour pruning.
    reference: utils/graph.py:40-87"""
    dev = _dev()
    keep = ops.knn_mask(torch.as_tensor(pwise_dist).to(dev, torch.float32),
                        torch.as_tensor(edge_ixs).to(dev, torch.int64), num_nodes, top_k_nns,
                        reciprocal_k_nns, symmetric_edges)
    return keep if use_cuda else keep.cpu()


def _col(det_df, name, dev):
    v = det_df[name]
    if torch.is_tensor(v):
        return v.to(dev).float()
    v = v.values if hasattr(v, 'values') else v
    return torch.as_tensor(np.asarray(v)).to(dev).float()       # float64 column -> fp32 (.float())


def compute_edge_feats_dict(edge_ixs, det_df, fps, use_cuda):
    """Dict of the five geometric features, each a FloatTensor [num_edges].
    reference: utils/graph.py:90-124"""
    dev = _dev()
    pairs = torch.as_tensor(edge_ixs).to(dev, torch.int64)
    attr, _ = ops.edge_feats_assemble(pairs, _col(det_df, 'frame', dev), _col(det_df, 'bb_height', dev),
                                      _col(det_df, 'bb_width', dev), _col(det_df, 'feet_x', dev),
                                      _col(det_df, 'feet_y', dev), fps, None)
    p = pairs.shape[1]
    names = ('secs_time_dists', 'norm_feet_x_dists', 'norm_feet_y_dists', 'bb_height_dists', 'bb_width_dists')
    out = {n: attr[:p, i].contiguous() for i, n in enumerate(names)}
    return out if use_cuda else {k: v.cpu() for k, v in out.items()}


def to_undirected_graph(mot_graph, attrs_to_update=('edge_preds', 'edge_labels')):
    """Keep one (i < j) copy of every directed edge pair of ``mot_graph.graph_obj`` (sorted by (i, j)) and
    average the attributes in ``attrs_to_update`` over the two copies.  Runs on the device the graph
    lives on.  reference: utils/graph.py:165-185"""
    go = mot_graph.graph_obj
    ei = go.edge_index
    e = ei.shape[1]
    half = e // 2
    # fast path: the layout every graph built here has -- [pairs with i < j sorted by (i, j) | the same pairs flipped]
    structured = e % 2 == 0 and e > 0 and bool(
        (ei[0, :half] < ei[1, :half]).all() and torch.equal(ei[:, half:], ei[:, :half].flip(0)))
    if structured:
        n = int(ei.max()) + 1
        key = ei[0, :half] * n + ei[1, :half]
        structured = bool((key[1:] > key[:-1]).all())
    if structured:
        go.edge_index = ei[:, :half].contiguous()
        for name in attrs_to_update:
            if hasattr(go, name):
                a = getattr(go, name)
                setattr(go, name, (a[:half] + a[half:]) / 2)
        return
    sorted_edges, _ = torch.sort(ei, dim=0)
    undirected, inverse = torch.unique(sorted_edges, return_inverse=True, dim=1)
    assert sorted_edges.shape[1] == 2 * undirected.shape[1], "Some edges were not duplicated"
    go.edge_index = undirected
    for name in attrs_to_update:
        if hasattr(go, name):
            a = getattr(go, name)
            s = torch.zeros(undirected.shape[1], dtype=a.dtype, device=a.device).index_add_(0, inverse, a)
            c = torch.zeros_like(s).index_add_(0, inverse, torch.ones_like(a))
            setattr(go, name, s / c.clamp(min=1))                       # scatter_mean


def to_lightweight_graph(mot_graph, attrs_to_del=('reid_emb_dists', 'x', 'edge_attr', 'edge_labels')):
    """Delete what inference no longer needs and prune edges predicted below 0.5.
    reference: utils/graph.py:187-207"""
    go = mot_graph.graph_obj
    go.num_nodes = go.num_nodes                                         # pin the count: ``x`` is deleted below (:194)
    go.node_names = torch.arange(go.num_nodes, device=go.edge_index.device)
    for name in attrs_to_del:
        if hasattr(go, name):
            delattr(go, name)
    keep = go.edge_preds >= 0.5
    go.edge_index = go.edge_index.T[keep].T
    go.edge_preds = go.edge_preds[keep]
