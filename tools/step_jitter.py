"""Development: per-step wall/device time of the bench step to find sporadic slow steps."""
import sys, os, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpntrackseg_b200 import synth
from mpntrackseg_b200.config import default_dataset_params, default_graph_model_params
from mpntrackseg_b200.data.mot_graph import build_graph_batch
from mpntrackseg_b200.models.mpn import MOTMPNet
dev = torch.device('cuda'); G = 16
ds = default_dataset_params(50, 15); mp = default_graph_model_params(12, 11)
P = synth.make_params(mp, seed=9, gain=1.25, core_only=True)
model = MOTMPNet(mp).to(dev).eval(); model.load_state_dict(P, strict=False)
wins = [synth.make_window(T=15, D=150, k=50, seed=g, node_dim=8) for g in range(G)]
ptr = [0]
for w in wins: ptr.append(ptr[-1] + w.N)
tab = {k: torch.cat([torch.from_numpy(synth.det_columns(w)[k]) for w in wins]).to(dev) for k in ('frame','bb_height','bb_width','feet_x','feet_y')}
tab['reid'] = torch.cat([w.reid for w in wins]).to(dev)
tab['x'] = [torch.randn(w.N, 2048, 8, 4, device=dev).abs_() for w in wins]
def step():
    b = build_graph_batch(tab, ptr, ds, 30.0, device=dev)
    with torch.no_grad(): return model.forward_batch(b)
ts = []
for i in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print('per-step ms:', ' '.join(f'{t:.1f}' for t in ts))
print('mem reserved GB', torch.cuda.memory_reserved() / 1e9, 'num_alloc_retries', torch.cuda.memory_stats().get('num_alloc_retries'), 'segments', torch.cuda.memory_stats().get('segment.all.allocated'))
