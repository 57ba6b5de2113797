"""Development check: where the batched builder (dense distance kernels) and get_knn_mask on the sequence graph's
edge-list distances keep different pairs, show the float64 gap at the k boundary of the rows involved."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpntrackseg_b200 import synth, ops
from mpntrackseg_b200.config import default_dataset_params
from mpntrackseg_b200.data.mot_graph import MOTGraph, build_graph_batch
from mpntrackseg_b200.utils.graph import get_knn_mask
dev = torch.device('cuda')
T, D, FPG, K = 45, 150, 15, 50
win = synth.make_window(T=T, D=D, k=K, seed=3, node_feats='pooled')
ds = default_dataset_params(K, FPG)
cols = synth.det_columns(win)
frame = win.frame.to(dev)
reid = win.reid.to(dev)
nper = D
tot_diff = 0
for engine in ('tc', 'fp32'):
    for t in range(T - FPG + 1):
        n0, n1 = t * nper, (t + FPG) * nper
        table = {c: torch.as_tensor(cols[c][n0:n1]).to(dev) for c in ('frame', 'bb_height', 'bb_width', 'feet_x', 'feet_y')}
        table['reid'], table['x'] = reid[n0:n1], torch.zeros(n1 - n0, 32, device=dev)
        b = build_graph_batch(table, [0, n1 - n0], ds, 30.0, max_frame_dist=FPG - 1, device=dev, engine=engine)
        P = b.pair_ptr[-1]
        got = set(map(tuple, b.edge_index[:, :P].T.tolist()))
        pairs = ops.time_valid_pairs(frame[n0:n1], FPG - 1)
        d = ops.pair_reid_dist(reid[n0:n1].contiguous(), pairs)
        ei = torch.cat((pairs, pairs.flip(0)), 1)
        keep = get_knn_mask(torch.cat((d, d)), ei, n1 - n0, K, True, reciprocal_k_nns=True, symmetric_edges=True)
        exp = set(map(tuple, pairs[:, keep[:pairs.shape[1]]].T.tolist()))
        diff = sorted(got ^ exp)
        tot_diff += len(diff)
        for (i, j) in diff[:3]:
            r64 = reid[n0:n1].double()
            for a, bb in ((i, j), (j, i)):
                dd = ((r64[a] - r64 + 1e-6) ** 2).sum(1).sqrt()
                dd[(frame[n0:n1] == frame[n0 + a])] = float('inf')
                s = torch.sort(dd).values
                print(f'{engine} window {t} pair ({i},{j}) row {a}: d={float(dd[bb]):.9f} k-th={float(s[K-1]):.9f} (k+1)-th={float(s[K]):.9f} '
                      f'rel gap={(float(s[K]) - float(s[K-1])) / float(s[K-1]):.2e}')
    print(engine, 'windows checked', T - FPG + 1, 'differing pairs in total', tot_diff)
