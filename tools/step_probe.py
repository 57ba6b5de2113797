"""Development: host-side time split of one bench step (graph build vs forward), device-resident inputs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpntrackseg_b200 import synth, ops
from mpntrackseg_b200.config import default_dataset_params, default_graph_model_params
from mpntrackseg_b200.data.mot_graph import build_window_graphs
from mpntrackseg_b200.models.mpn import MOTMPNet
dev = torch.device('cuda'); G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ds = default_dataset_params(50, 15); mp = default_graph_model_params(12, 11)
P = synth.make_params(mp, seed=9, gain=1.25, core_only=True)
model = MOTMPNet(mp).to(dev).eval(); model.load_state_dict(P, strict=False)
inputs = []
for g in range(G):
    w = synth.make_window(T=15, D=150, k=50, seed=g, node_dim=8)
    d = {k: torch.from_numpy(v).to(dev) for k, v in synth.det_columns(w).items()}
    d['reid'] = w.reid.to(dev); d['x'] = torch.randn(w.N, 2048, device=dev).abs_()
    inputs.append(d)
def t(fn, name, n=5):
    for _ in range(3): r = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); print(f'{name}: {(time.perf_counter()-t0)/n*1e3:.2f} ms'); return r
batch = t(lambda: build_window_graphs(inputs, ds, 30.0, device=dev), 'build_window_graphs')
frame = torch.cat([d['frame'].to(torch.int64) for d in inputs]); reid = torch.cat([d['reid'] for d in inputs])
ptr = batch.node_ptr
t(lambda: ops.knn_graph_pairs(frame, ptr, reid, 50, True, -1), '  knn_graph_pairs')
t(lambda: ops.knn_graph_pairs(frame, ptr, reid, 50, True, -1, engine='fp32'), '  knn_graph_pairs fp32')
with torch.no_grad():
    t(lambda: model.forward_batch(batch), 'forward_batch')
    lay = t(lambda: ops.edge_layout(batch.edge_index, batch.num_nodes), '  edge_layout')
    x0 = t(lambda: model.encode_nodes_list(batch.xs), '  encode_nodes_list')
    e0 = t(lambda: model.encode_edges(batch.edge_attr, lay), '  encode_edges')
    cw, keep = model.core_weights()
    t(lambda: ops.mp_forward(cw, lay, x0, e0, 12, 2, engine='tc'), '  mp_forward tc')
