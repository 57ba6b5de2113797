"""Measurement for SURVEY.md section 8 row f1: whole-sequence sliding-window evaluation
(MPNTracker._evaluate_graph_in_batches), batched B200 schedule vs the reference's window-by-window schedule,
both on the GPU.  Usage: python tools/track_seq_bench.py [T] [D] [frames_per_graph] [k]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mpntrackseg_b200 import synth
from mpntrackseg_b200.config import default_dataset_params, default_graph_model_params
from mpntrackseg_b200.data.mot_graph import MOTGraph
from mpntrackseg_b200.models.mpn import MOTMPNet
from mpntrackseg_b200.tracker import MPNTracker

T = int(sys.argv[1]) if len(sys.argv) > 1 else 45
D = int(sys.argv[2]) if len(sys.argv) > 2 else 150
FPG = int(sys.argv[3]) if len(sys.argv) > 3 else 15
K = int(sys.argv[4]) if len(sys.argv) > 4 else 50
dev = torch.device('cuda')


class Full(object):
    pass


def full_graph(win, ds):
    mg = MOTGraph.from_tensors(synth.det_columns(win), win.reid, win.x.to(dev), None, {'fps': win.fps}, ds, inference_mode=True,
                  max_frame_dist=FPG - 1)
    mg.construct_graph_object()
    mg.frames = sorted(set(win.frame.tolist()))
    mg.frames_per_graph = FPG
    return mg


def main():
    win = synth.make_window(T=T, D=D, k=K, seed=3, node_feats='full')
    ds = default_dataset_params(K, FPG)
    mp = default_graph_model_params(12, 11)
    model = MOTMPNet(mp).to(dev).eval()
    model.load_state_dict(synth.make_params(mp, seed=9, gain=1.25, core_only=True), strict=False)
    tr = MPNTracker(graph_model=model, eval_params={'set_pruned_edges_to_inactive': True}, dataset_params=ds,
                    window_batch=16)
    res = {}
    for name in ('batched_rebuild', 'batched', 'window_by_window') * 2:
        tr.full_graph = full_graph(win, ds)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if name == 'batched_rebuild':
            tr._evaluate_batched_rebuild()
        elif name == 'batched':
            tr._evaluate_batched()
        else:
            tr._evaluate_window_by_window(None)
        torch.cuda.synchronize()
        res[name] = (time.perf_counter() - t0, tr.full_graph.graph_obj.edge_preds)
    go = tr.full_graph.graph_obj
    nwin = T - FPG + 1
    diff = max(float((res[a][1] - res['window_by_window'][1]).abs().max()) for a in ('batched', 'batched_rebuild'))
    print(f'sequence: T={T} frames, N={win.N} nodes, E_full={go.num_edges} directed candidate edges, '
          f'{nwin} windows of {FPG} frames, k={K}')
    for name in ('batched_rebuild', 'batched', 'window_by_window'):
        print(f'  {name:17s}: {res[name][0] * 1e3:8.1f} ms  ({nwin / res[name][0]:7.1f} windows/s)')
    print(f'  max |edge_pred difference| between the schedules: {diff:.2e}')
    for a in ('batched', 'batched_rebuild'):
        d = (res[a][1] - res['window_by_window'][1]).abs()
        print(f'    {a}: {int((d > 1e-3).sum())} of {d.numel()} directed edges differ by more than 1e-3 '
              f'(near-ties at the k boundary are resolved by the last ulp of the distance, whose summation order differs '
              f'between the edge-list and the dense kernels)')


if __name__ == '__main__':
    main()
