"""CPU oracle for what follows the hot path: constraint statistics, greedy rounding, identities.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain numpy + Python loops, written from the reference's behaviour:

* ``constr_satisfaction_rate`` -- src/mot_neural_solver/utils/evaluation.py:370-414
* ``greedy_project``           -- src/mot_neural_solver/tracker/projectors.py:11-67 (GreedyProjector.project)
* ``connected_components``     -- src/mot_neural_solver/tracker/mpn_tracker.py:231-248 (scipy csgraph, undirected)

Pinned by tests/test_oracle_golden.py against tests/golden/rounding.npz (outputs of the reference's own functions).
"""
import numpy as np


def constr_satisfaction_rate(edge_index, num_nodes, edges_out, undirected_edges=True):
    """evaluation.py:370-414.  edges_out is the BINARISED edge vector.  Returns (rate, flow_in, flow_out).

    With both directions listed (undirected_edges) every pair is ordered (min, max) first and counted half
    (:391-397); a constraint exists for every node that is the source (resp. target) of at least one listed edge
    (:408-409), whatever the edge's value."""
    ei = np.asarray(edge_index, dtype=np.int64)
    v = np.asarray(edges_out, dtype=np.float32)
    if undirected_edges:
        src, dst, div = np.minimum(ei[0], ei[1]), np.maximum(ei[0], ei[1]), np.float32(2.0)
    else:
        src, dst, div = ei[0], ei[1], np.float32(1.0)
    # fp32 accumulation of 0/1 values is exact, so the summation order of scatter_add does not matter
    flow_out = np.bincount(src, weights=v, minlength=num_nodes).astype(np.float32) / div
    flow_in = np.bincount(dst, weights=v, minlength=num_nodes).astype(np.float32) / div
    violated = np.float32(int((flow_in > 1).sum()) + int((flow_out > 1).sum()))
    num_constraints = len(np.unique(src)) + len(np.unique(dst))
    rate = np.float32(1) - violated / np.float32(num_constraints)        # the reference divides two fp32 tensors
    return float(rate), flow_in, flow_out


def greedy_project(edge_index, edge_preds, num_nodes):
    """projectors.py:19-67.  Returns (rounded fp32 0/1 vector, constraint rate of the plain rounding).

    Edges are listed once (i < j).  Violated flow-out constraints are handled before flow-in ones (the descending sort
    on the type column, :38-40); constraints of one type touch disjoint edge sets, so their mutual order is free.  A
    constraint that is still violated keeps the incident edge with the largest pred * current rounding value -- Python's
    max returns the FIRST maximum in ascending edge order (:56-57) -- and switches the others off."""
    ei = np.asarray(edge_index, dtype=np.int64)
    preds = np.asarray(edge_preds, dtype=np.float32)
    rounded = (preds > 0.5).astype(np.float32)
    rate, flow_in, flow_out = constr_satisfaction_rate(ei, num_nodes, rounded, undirected_edges=False)
    todo = [(int(n), 1) for n in np.nonzero(flow_out > 1)[0]] + [(int(n), 0) for n in np.nonzero(flow_in > 1)[0]]
    for node, kind in todo:
        incident = np.nonzero(ei[1 if kind == 0 else 0] == node)[0]
        if rounded[incident].sum() > 1:
            score = preds[incident] * rounded[incident]
            best = incident[int(np.argmax(score))]                       # argmax: first maximum, like max()
            rounded[incident] = 0
            rounded[best] = 1
    assert np.bincount(ei[1], weights=rounded, minlength=num_nodes).max(initial=0) <= 1
    assert np.bincount(ei[0], weights=rounded, minlength=num_nodes).max(initial=0) <= 1
    return rounded, rate


def connected_components(edge_index, edge_preds, num_nodes):
    """mpn_tracker.py:231-242: components of the edges whose value is exactly 1, undirected.  Labels are numbered in
    the order scipy meets the components when it walks the nodes 0..n-1, i.e. by each component's smallest node."""
    ei = np.asarray(edge_index, dtype=np.int64)
    on = np.asarray(edge_preds) == 1
    parent = list(range(num_nodes))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in zip(ei[0][on].tolist(), ei[1][on].tolist()):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)                            # the root is the smallest node of the component
    labels = np.empty(num_nodes, dtype=np.int64)
    next_label = 0
    for node in range(num_nodes):
        root = find(node)
        if root == node:                                                 # first (smallest) node of a new component
            labels[node] = next_label
            next_label += 1
        else:
            labels[node] = labels[root]
    return next_label, labels
