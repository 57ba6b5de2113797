"""Oracle (test infrastructure): the tracker's sliding-window evaluation on CPU.

Restates tracker/mpn_tracker.py:96-141 (``_predict_edges_and_masks``), :143-210
(``_evaluate_graph_in_batches``) and utils/graph.py:165-207 (``to_undirected_graph``,
``to_lightweight_graph``) of the reference in CPU PyTorch, around the oracle's own graph / network
functions.  The reference module itself cannot be imported in this image (torch_geometric,
pycocotools missing); ``tests/golden/make_golden.py`` runs the same loop around the IMPORTED
reference functions to pin this file (fixture ``tracker_sequence.npz``).
See ``oracle/__init__.py`` for who may import this.
"""
import torch

from . import graph_ref, mpn_ref


def predict_window_edges(P, model_params, dataset_params, eval_params, x, edge_index, edge_attr, reid_emb_dists):
    """KNN-prune one window, run the network, scatter sigmoid(last logits) back to the window's
    unpruned edge list.  Returns (edge_preds [E'], pred_mask [E'] bool).
    reference: tracker/mpn_tracker.py:107-141"""
    ei, ea, keep = graph_ref.prune_window(edge_index, edge_attr, reid_emb_dists, x.shape[0], dataset_params)
    out = mpn_ref.mpn_forward(P, model_params, x, ei, ea)
    preds = mpn_ref.window_edge_preds(out['classified_edges'], keep)
    if eval_params['set_pruned_edges_to_inactive']:                      # mpn_tracker.py:137-138
        return preds, torch.ones_like(keep)
    return preds, keep


def evaluate_graph_in_batches(P, model_params, dataset_params, eval_params, frame_num, x, edge_index, edge_attr,
                              reid_emb_dists, frames_per_graph):
    """Average per-edge predictions over all sliding windows of ``frames_per_graph`` frames.
    Returns the directed per-edge predictions [E_full] (before the undirected merge).
    reference: tracker/mpn_tracker.py:153-205 (edge part)"""
    frame_num = torch.as_tensor(frame_num).view(-1)
    all_frames = torch.unique(frame_num)                                 # sorted
    total = torch.zeros(edge_index.shape[1])
    count = torch.zeros(edge_index.shape[1])
    node_names = torch.arange(x.shape[0])
    for start, end in zip(all_frames, all_frames[frames_per_graph - 1:]):
        nodes_mask = (start <= frame_num) & (frame_num <= end)           # :171
        edges_mask = nodes_mask[edge_index[0]] & nodes_mask[edge_index[1]]
        sub_ei = edge_index.T[edges_mask].T - node_names[nodes_mask][0]  # :179
        preds, pred_mask = predict_window_edges(P, model_params, dataset_params, eval_params, x[nodes_mask], sub_ei,
                                                edge_attr[edges_mask], reid_emb_dists[edges_mask])
        total[edges_mask] += preds                                       # :195
        ids = torch.where(edges_mask)[0][pred_mask]
        count[ids] += 1                                                  # :197
    final = total / count                                                # :203
    final[torch.isnan(final)] = 0                                        # :204
    return final


def to_undirected(edge_index, edge_attrs):
    """Keep one (i < j) copy of every edge pair, sorted by (i, j); average each attribute over the
    two directed copies.  reference: utils/graph.py:165-185"""
    sorted_edges, _ = torch.sort(edge_index, dim=0)
    undirected, inverse = torch.unique(sorted_edges, return_inverse=True, dim=1)
    assert sorted_edges.shape[1] == 2 * undirected.shape[1], 'Some edges were not duplicated'
    out = []
    for a in edge_attrs:
        s = torch.zeros(undirected.shape[1], dtype=a.dtype).index_add_(0, inverse, a)
        c = torch.zeros(undirected.shape[1], dtype=a.dtype).index_add_(0, inverse, torch.ones_like(a))
        out.append(s / c.clamp(min=1))
    return undirected, out


def to_lightweight(edge_index, edge_preds):
    """Drop edges whose averaged prediction is below 0.5.  reference: utils/graph.py:204-207"""
    keep = edge_preds >= 0.5
    return edge_index.T[keep].T, edge_preds[keep]
