"""Pins the CPU oracle (oracle/) against outputs of the unmodified reference
(tests/golden/*.npz, written by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from cases import CASES, load_case
from mpntrackseg_b200 import synth
from oracle import graph_ref, mpn_ref

GRAPH_CASES = list(CASES)


@pytest.mark.parametrize('name', GRAPH_CASES)
def test_graph_build_matches_reference(name):
    c = load_case(name)
    win, gold = c['win'], c['gold']
    g = graph_ref.build_graph(win.frame, win.reid, synth.det_columns(win), win.fps, c['ds'],
                              inference_mode=False, max_frame_dist=c['max_frame_dist'])
    pairs = graph_ref.time_valid_pairs(win.frame, c['max_frame_dist'])
    assert pairs.shape[1] == int(gold['n_candidates'])
    assert np.array_equal(g['edge_index'].numpy(), gold['edge_index'].astype(np.int64))
    # same torch ops on the same CPU: bit-exact features
    assert np.array_equal(g['edge_attr'].numpy(), gold['edge_attr'])


@pytest.mark.parametrize('name', GRAPH_CASES)
def test_core_forward_matches_reference(name):
    c = load_case(name)
    win, gold = c['win'], c['gold']
    ei = torch.from_numpy(gold['edge_index'].astype(np.int64))
    ea = torch.from_numpy(gold['edge_attr'])
    with torch.no_grad():
        out = mpn_ref.mpn_forward(c['P'], c['mp'], win.x, ei, ea, return_state=True)
    logits = torch.stack([t.view(-1) for t in out['classified_edges']]).numpy()
    assert logits.shape == gold['logits'].shape
    # identical op sequence -> expect (near) bit equality; allow 1e-5 for BLAS blocking
    np.testing.assert_allclose(logits, gold['logits'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out['edge_state'].numpy(), gold['edge_state'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out['node_state'].numpy(), gold['node_state'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('name', ['tiny_full'])
def test_full_forward_with_mask_branch(name):
    c = load_case(name)
    win, gold = c['win'], c['gold']
    ei = torch.from_numpy(gold['edge_index'].astype(np.int64))
    ea = torch.from_numpy(gold['edge_attr'])
    with torch.no_grad():
        out = mpn_ref.mpn_forward(c['P'], c['mp'], win.x, ei, ea, x_ext=win.x_ext)
    logits = torch.stack([t.view(-1) for t in out['classified_edges']]).numpy()
    np.testing.assert_allclose(logits, gold['logits'], rtol=1e-5, atol=1e-5)
    assert len(out['mask_predictions']) == c['mp']['num_class_steps']
    m = out['mask_predictions'][-1]
    assert tuple(m.shape) == (win.N, 1, 56, 56)
    np.testing.assert_allclose(m[:, 0, ::7, ::7].numpy(), gold['mask_last_sample'], rtol=1e-4, atol=1e-4)
    means = np.array([float(t.double().mean()) for t in out['mask_predictions']])
    np.testing.assert_allclose(means, gold['mask_step_means'], rtol=1e-4, atol=1e-5)


def test_tracker_window_glue():
    c = load_case('tracker_window')
    win, gold, ds = c['win'], c['gold'], c['ds']
    fpg = ds['frames_per_graph']
    g = graph_ref.build_graph(win.frame, win.reid, synth.det_columns(win), win.fps, ds,
                              inference_mode=True, max_frame_dist=fpg - 1)
    assert np.array_equal(g['edge_index'].numpy(), gold['full_edge_index'].astype(np.int64))
    assert np.array_equal(g['edge_attr'].numpy(), gold['full_edge_attr'])
    assert np.array_equal(g['reid_emb_dists'].numpy(), gold['full_dists'])
    start, end = (int(v) for v in gold['window'])
    nodes = (win.frame >= start) & (win.frame <= end)
    edges = nodes[g['edge_index'][0]] & nodes[g['edge_index'][1]]
    first = int(torch.nonzero(nodes)[0])
    sub_ei = g['edge_index'][:, edges] - first
    ei, ea, keep = graph_ref.prune_window(sub_ei, g['edge_attr'][edges], g['reid_emb_dists'][edges],
                                          int(nodes.sum()), ds)
    assert np.array_equal(keep.numpy(), gold['keep'])
    with torch.no_grad():
        out = mpn_ref.mpn_forward(c['P'], c['mp'], win.x[nodes], ei, ea)
    preds = mpn_ref.window_edge_preds(out['classified_edges'], keep)
    np.testing.assert_allclose(preds.numpy(), gold['edge_preds'], rtol=1e-5, atol=1e-6)


def test_tracker_sequence_sliding_windows():
    """mpn_tracker.py:153-207 restatement (all windows, per-edge averaging, undirected merge, pruning at 0.5)
    against the loop run around the imported reference functions (fixture tracker_sequence.npz)."""
    from oracle import tracker_ref
    c = load_case('tracker_window')
    seq = dict(np.load(os.path.join(os.path.dirname(__file__), 'golden', 'tracker_sequence.npz')))
    assert str(seq['param_checksum']) == str(c['gold']['param_checksum'])
    win, ds = c['win'], c['ds']
    fpg = ds['frames_per_graph']
    g = graph_ref.build_graph(win.frame, win.reid, synth.det_columns(win), win.fps, ds,
                              inference_mode=True, max_frame_dist=fpg - 1)
    for tag, inactive in (('knn', False), ('inactive', True)):
        with torch.no_grad():
            final = tracker_ref.evaluate_graph_in_batches(
                c['P'], c['mp'], ds, {'set_pruned_edges_to_inactive': inactive}, win.frame, win.x, g['edge_index'],
                g['edge_attr'], g['reid_emb_dists'], fpg)
        np.testing.assert_allclose(final.numpy(), seq[f'directed_preds_{tag}'], rtol=1e-5, atol=1e-6)
        und_ei, (und_p,) = tracker_ref.to_undirected(g['edge_index'], [final])
        assert np.array_equal(und_ei.numpy(), seq[f'undirected_edge_index_{tag}'].astype(np.int64))
        np.testing.assert_allclose(und_p.numpy(), seq[f'undirected_preds_{tag}'], rtol=1e-5, atol=1e-6)
        # pruning at 0.5 from the fixture's own averaged predictions (no threshold noise)
        li, lp = tracker_ref.to_lightweight(und_ei, torch.from_numpy(seq[f'undirected_preds_{tag}']))
        assert np.array_equal(li.numpy(), seq[f'light_edge_index_{tag}'].astype(np.int64))
        assert np.array_equal(lp.numpy(), seq[f'light_preds_{tag}'])


def test_edge_labels_against_reference_formulation_and_brute_force():
    """data/mot_graph.py:223-262: oracle vs the fixture made with the reference's scatter_min data flow, and vs
    the definition (closest same-identity partner by node index, per direction)."""
    gold = dict(np.load(os.path.join(os.path.dirname(__file__), 'golden', 'edge_labels.npz')))
    ei, ids = torch.from_numpy(gold['edge_index'].astype(np.int64)), torch.from_numpy(gold['ids'])
    assert np.array_equal(graph_ref.assign_edge_labels(ei, ids, 'all').numpy(), gold['labels_all'])
    got = graph_ref.assign_edge_labels(ei, ids, 'closest')
    assert np.array_equal(got.numpy(), gold['labels_closest'])
    nbrs = {}
    for e, (r, c) in enumerate(ei.T.tolist()):
        if ids[r] == ids[c] and ids[r] != -1:
            nbrs.setdefault((r, c > r), []).append(c)
    exp = torch.zeros(ei.shape[1])
    for e, (r, c) in enumerate(ei.T.tolist()):
        cand = nbrs.get((r, c > r), [])
        if ids[r] == ids[c] and ids[r] != -1 and c == (min(cand) if c > r else max(cand)):
            exp[e] = 1
    assert torch.equal(got, exp)
    assert 0 < exp.sum() < gold['labels_all'].sum()


def test_knn_mask_independent_restatement():
    """The dense-argsort restatement agrees with a per-node sort formulation."""
    win = synth.make_window(T=7, D=9, k=8, seed=5)
    pairs = graph_ref.time_valid_pairs(win.frame)
    d = graph_ref.pair_reid_dist(win.reid, pairs)
    for recip in (True, False):
        got = graph_ref.knn_keep_mask(d, pairs, win.N, 8, recip, symmetric_edges=False)
        nbr = [[] for _ in range(win.N)]
        for e in range(pairs.shape[1]):
            i, j = int(pairs[0, e]), int(pairs[1, e])
            nbr[i].append((float(d[e]), j))
            nbr[j].append((float(d[e]), i))
        top = [set(j for _, j in sorted(l)[:8]) for l in nbr]
        exp = []
        for e in range(pairs.shape[1]):
            i, j = int(pairs[0, e]), int(pairs[1, e])
            a, b = j in top[i], i in top[j]
            exp.append((a and b) if recip else (a or b))
        assert got.tolist() == exp


def test_segment_softmax_sums_to_one_and_handles_gaps():
    src = torch.tensor([[0.3], [2.0], [-1.0], [5.0]])
    idx = torch.tensor([0, 0, 3, 3])
    w = mpn_ref.segment_softmax(src, idx)
    s = mpn_ref.segment_add(w, idx, 4).view(-1)
    assert torch.allclose(s, torch.tensor([1.0, 0.0, 0.0, 1.0]), atol=1e-6)


def test_loss_zero_positives_and_weighting():
    logits = [torch.tensor([[0.2], [-0.4], [1.5]])]
    assert float(mpn_ref.weighted_bce_loss(logits, torch.zeros(3))) > 0
    lab = torch.tensor([1.0, 0.0, 0.0])
    l = mpn_ref.weighted_bce_loss(logits, lab)
    x = logits[0].view(-1)
    manual = (2.0 * lab * torch.nn.functional.softplus(-x) + (1 - lab) * torch.nn.functional.softplus(x)).mean()
    assert torch.allclose(l, manual, atol=1e-6)


# ------------------------------------------------------------------ the large fixtures (BASELINE.json configs[4], > 5,120 nodes)
def test_config5_graph_and_forward_match_reference():
    """configs[4] (15 x 300 detections, k = 100; N = 4,500, E = 179 k): the oracle's edge list, features and logits against
    the reference's own (tests/golden/config5.npz keeps a strided feature sample, two logit rows and the step means)."""
    from cases import load_big_case
    c = load_big_case('config5')
    win, ds, gold = c['win'], c['ds'], c['gold']
    g = graph_ref.build_graph(win.frame, win.reid, synth.det_columns(win), win.fps, ds, inference_mode=False, max_frame_dist='max')
    assert graph_ref.time_valid_pairs(win.frame, 'max').shape[1] == int(gold['n_candidates'])
    assert np.array_equal(g['edge_index'].numpy(), gold['edge_index'].astype(np.int64))
    assert np.array_equal(g['edge_attr'][::16].numpy(), gold['edge_attr_sample'])
    with torch.no_grad():
        out = mpn_ref.mpn_forward(c['P'], c['mp'], win.x, g['edge_index'], g['edge_attr'])
    logits = torch.stack([t.view(-1) for t in out['classified_edges']])
    # node state ~5e4 after 12 sum aggregations: fp32 summation-order noise is ~1e-4 relative on these logits
    for row, key in ((-1, 'logits_last'), (0, 'logits_first')):
        ref = gold[key]
        assert np.all(np.abs(logits[row].numpy() - ref) <= 1e-3 * np.maximum(1.0, np.abs(ref)))
    np.testing.assert_allclose(logits.double().mean(dim=1).numpy(), gold['logits_step_means'], rtol=1e-4, atol=1e-4)


def test_big_window_pairs_match_reference():
    """A 5,400-node window (tests/golden/big_window.npz): kept pairs and a sample of their distances."""
    from cases import load_big_case
    c = load_big_case('big_window')
    win, ds, gold = c['win'], c['ds'], c['gold']
    g = graph_ref.build_graph(win.frame, win.reid, synth.det_columns(win), win.fps, ds, inference_mode=False, max_frame_dist='max')
    p = g['edge_index'].shape[1] // 2
    assert np.array_equal(g['edge_index'][:, :p].numpy(), gold['pairs'].astype(np.int64))
    assert np.array_equal(g['edge_attr'][:p:8, 5].numpy(), gold['reid_dist_sample'])


# ------------------------------------------------------------------ rounding + identities
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_rounding_oracle_matches_reference(tag):
    """oracle/rounding_ref.py against the reference's compute_constr_satisfaction_rate / GreedyProjector and scipy's
    connected_components (tests/golden/rounding.npz): integer work, bit-exact."""
    from oracle import rounding_ref
    gold = dict(np.load(os.path.join(os.path.dirname(__file__), 'golden', 'rounding.npz')))
    n = int(gold[f'n_{tag}'])
    ei, preds = gold[f'edge_index_{tag}'].astype(np.int64), gold[f'preds_{tag}']
    r0 = (preds > 0.5).astype(np.float32)
    both = np.concatenate((ei, ei[::-1]), axis=1)
    rate_u, fin, fout = rounding_ref.constr_satisfaction_rate(both, n, np.concatenate((r0, r0)), undirected_edges=True)
    assert rate_u == float(gold[f'rate_undirected_{tag}'])
    assert np.array_equal(fin, gold[f'flow_in_{tag}']) and np.array_equal(fout, gold[f'flow_out_{tag}'])
    rounded, rate = rounding_ref.greedy_project(ei, preds, n)
    assert rate == float(gold[f'rate_{tag}'])
    assert np.array_equal(rounded, gold[f'round_{tag}'])
    assert int((rounded > r0).sum()) == 0                                    # the projection only switches edges off
    ncomp, labels = rounding_ref.connected_components(ei, rounded, n)
    assert ncomp == int(gold[f'ncomp_{tag}']) and np.array_equal(labels, gold[f'labels_{tag}'])


def test_rounding_oracle_components_equal_scipy_on_random_graphs():
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    from oracle import rounding_ref
    rng = np.random.default_rng(5)
    for n, e in ((1, 0), (7, 3), (60, 45), (300, 500)):
        ei = rng.integers(0, n, size=(2, e))
        on = (rng.random(e) < 0.6).astype(np.float32)
        ncomp, labels = rounding_ref.connected_components(ei, on, n)
        m = on == 1
        ref_n, ref_l = connected_components(csr_matrix((np.ones(int(m.sum()), dtype=int), tuple(ei[:, m])), shape=(n, n)),
                                            directed=False, return_labels=True)
        assert ncomp == ref_n and np.array_equal(labels, ref_l)
