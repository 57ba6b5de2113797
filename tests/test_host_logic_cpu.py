"""Host-side logic that needs no GPU: the undirected merge / pruning of the sequence graph and the tracker
driver's window bookkeeping (pure index arithmetic in torch), against the reference-generated fixture."""
import os

import numpy as np
import torch

from mpntrackseg_b200.data.mot_graph import Graph
from mpntrackseg_b200.tracker.mpn_tracker import MPNTracker
from mpntrackseg_b200.utils.graph import to_lightweight_graph, to_undirected_graph

GOLD = dict(np.load(os.path.join(os.path.dirname(__file__), 'golden', 'tracker_sequence.npz')))


class _Full(object):
    pass


def _directed(tag):
    und = GOLD[f'undirected_edge_index_{tag}'].astype(np.int64)
    ei = torch.from_numpy(np.concatenate((und, und[::-1]), axis=1))
    return ei, torch.from_numpy(GOLD[f'directed_preds_{tag}'])


def test_to_undirected_and_lightweight_match_reference_fixture():
    """utils/graph.py:165-207 on the structured layout (fast path) and on a shuffled edge list (sort + unique path)."""
    for tag in ('knn', 'inactive'):
        ei, p = _directed(tag)
        n = int(ei.max()) + 1
        perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(5))
        for e_in, p_in in ((ei, p), (ei[:, perm], p[perm])):
            full = _Full()
            full.graph_obj = Graph(x=torch.zeros(n, 1), edge_index=e_in.clone(), edge_preds=p_in.clone(),
                                   edge_attr=torch.zeros(e_in.shape[1], 6))
            to_undirected_graph(full, attrs_to_update=('edge_preds', 'edge_labels'))
            go = full.graph_obj
            assert np.array_equal(go.edge_index.numpy(), GOLD[f'undirected_edge_index_{tag}'].astype(np.int64))
            np.testing.assert_allclose(go.edge_preds.numpy(), GOLD[f'undirected_preds_{tag}'], rtol=1e-6, atol=1e-7)
            go.edge_preds = torch.from_numpy(GOLD[f'undirected_preds_{tag}'])        # no threshold noise below
            to_lightweight_graph(full)
            assert np.array_equal(go.edge_index.numpy(), GOLD[f'light_edge_index_{tag}'].astype(np.int64))
            assert np.array_equal(go.edge_preds.numpy(), GOLD[f'light_preds_{tag}'])
            assert not hasattr(go, 'x') and not hasattr(go, 'edge_attr') and go.num_nodes == n
            assert torch.equal(go.node_names, torch.arange(n))


def test_structured_layout_detection():
    ei, _ = _directed('knn')
    go = Graph(x=torch.zeros(int(ei.max()) + 1, 1), edge_index=ei)
    assert MPNTracker._structured(go)
    half = ei.shape[1] // 2
    swapped = ei.clone()
    swapped[:, [0, 1]] = swapped[:, [1, 0]]                                           # first half no longer sorted
    assert not MPNTracker._structured(Graph(x=go.x, edge_index=swapped))
    assert not MPNTracker._structured(Graph(x=go.x, edge_index=ei[:, :-1]))           # odd number of edges
    gap = torch.cat((ei[:, :half][:, [0, 2]], ei[:, :half][:, [0, 2]].flip(0)), 1)    # a row with a hole in its partner range
    if int(gap[0, 0]) == int(gap[0, 1]) and int(gap[1, 1]) != int(gap[1, 0]) + 1:
        assert not MPNTracker._structured(Graph(x=go.x, edge_index=gap))


def test_window_ranges_and_prediction_counts():
    """Sliding windows (mpn_tracker.py:167-173) as contiguous node ranges, and the closed form of how many windows
    contain both endpoints of an edge, against the reference's mask formulation."""
    frames_per_node = torch.tensor([3, 3, 4, 6, 6, 6, 7, 9, 9, 10, 12, 12, 13])
    all_frames = sorted(set(frames_per_node.tolist()))
    fpg = 3
    tr = MPNTracker(graph_model=None, eval_params={'set_pruned_edges_to_inactive': True}, dataset_params={})
    tr.full_graph = _Full()
    tr.full_graph.frames, tr.full_graph.frames_per_graph = all_frames, fpg
    windows = tr._windows()
    assert len(windows) == len(all_frames) - fpg + 1
    n0s, n1s = tr._window_node_ranges(frames_per_node)
    n = frames_per_node.numel()
    pairs = torch.tensor([(i, j) for i in range(n) for j in range(i + 1, n) if frames_per_node[i] != frames_per_node[j]]).T
    brute = torch.zeros(pairs.shape[1])
    for (start, end), n0, n1 in zip(windows, n0s, n1s):
        mask = (int(start) <= frames_per_node) & (frames_per_node <= int(end))
        assert torch.equal(torch.nonzero(mask).view(-1), torch.arange(n0, n1))
        brute += (mask[pairs[0]] & mask[pairs[1]]).float()
    fpos = torch.searchsorted(torch.tensor(all_frames), frames_per_node)
    fi, fj = fpos[pairs[0]], fpos[pairs[1]]
    nwin = len(windows)
    closed = (torch.minimum(fi, torch.full_like(fi, nwin - 1)) - (fj - fpg + 1).clamp(min=0) + 1).clamp(min=0).float()
    assert torch.equal(closed, brute)
