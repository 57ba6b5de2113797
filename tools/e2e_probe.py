"""Breakdown of the end-to-end step (development probe, not part of the product)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpntrackseg_b200 import synth
from mpntrackseg_b200.config import default_dataset_params, default_graph_model_params
from mpntrackseg_b200.data.mot_graph import MOTGraph
from mpntrackseg_b200.models.mpn import MOTMPNet

dev = torch.device('cuda')
ds = default_dataset_params(50, 15)
mp = default_graph_model_params(12, 11)
P = synth.make_params(mp, seed=9, gain=1.25, core_only=True)
model = MOTMPNet(mp).to(dev).eval(); model.load_state_dict(P, strict=False)
G = 4
wins = [synth.make_window(T=15, D=150, k=50, seed=g, node_dim=8) for g in range(G)]
host = []
for w in wins:
    x = torch.randn(w.N, 2048, 8, 4).abs_()
    cols = {k: torch.from_numpy(v) for k, v in synth.det_columns(w).items()}
    host.append(dict(x=x.pin_memory(), reid=w.reid.pin_memory(), **{k: v.pin_memory() for k, v in cols.items()}))
print('pinned', host[0]['x'].is_pinned())

def t(fn, name, n=3):
    for _ in range(2): r = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); print(f'{name}: {(time.perf_counter()-t0)/n*1e3:.2f} ms'); return r

inputs = t(lambda: [{k: v.to(dev, non_blocking=True) for k, v in h.items()} for h in host], 'h2d')
graphs = t(lambda: [MOTGraph.from_tensors(d, d['reid'], d['x'], None, {'fps': 30.0}, ds).construct_graph_object() for d in inputs], 'graph build')
with torch.no_grad():
    x0 = t(lambda: [model.encode_nodes(g.x) for g in graphs], 'encode nodes')
    outs = t(lambda: model.forward_batch(graphs), 'forward_batch')
    t(lambda: [o['classified_edges'][-1].cpu() for o in outs], 'd2h')
