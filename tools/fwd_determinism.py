"""Development check: repeated forwards of one dense-crowd batch (D=300, k=150) -- bitwise equal? tc vs fp32?"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpntrackseg_b200 import synth
from mpntrackseg_b200.config import default_dataset_params, default_graph_model_params
from mpntrackseg_b200.data.mot_graph import build_window_graphs
from mpntrackseg_b200.models.mpn import MOTMPNet
dev = torch.device('cuda:0')
G, D, K = int(os.environ.get('G', 4)), int(os.environ.get('D', 300)), int(os.environ.get('K', 150))
ds = default_dataset_params(top_k_nns=K, frames_per_graph=15)
mp = default_graph_model_params(12, 11)
P = synth.make_params(mp, seed=9, gain=1.25, core_only=True)
model = MOTMPNet(mp).to(dev).eval(); model.load_state_dict(P, strict=False)
gen = torch.Generator(device=dev).manual_seed(0)
tabs = []
for s in range(G):
    w = synth.make_window(T=15, D=D, k=K, seed=s, node_feats='pooled', node_dim=8, min_gap=0)
    t = {k: torch.from_numpy(v) for k, v in synth.det_columns(w).items()}
    t['reid'] = w.reid; t['x'] = torch.randn((w.N, 2048), generator=gen, device=dev).abs_()
    tabs.append(t)
batch = build_window_graphs(tabs, ds, 30.0, device=dev)
outs = {}
for eng in ('tc', 'tc', 'auto', 'fp32'):
    model.engine = eng
    junk = torch.randn(32 << 20, device=dev); del junk
    try:
        with warnings.catch_warnings(record=True) as wl:
            warnings.simplefilter('always')
            with torch.no_grad():
                o = model.forward_batch(batch).logits[-1].clone()
        print(eng, 'ok absmax', float(o.abs().max()), 'warnings', [str(x.message)[:60] for x in wl])
        outs.setdefault(eng, []).append(o)
    except OverflowError as e:
        print(eng, 'OverflowError', e)
if len(outs.get('tc', [])) == 2:
    print('tc run1 == run2:', torch.equal(outs['tc'][0], outs['tc'][1]), float((outs['tc'][0] - outs['tc'][1]).abs().max()))
if 'tc' in outs and 'fp32' in outs:
    a, b = outs['tc'][0], outs['fp32'][0]
    print('tc vs fp32 max abs', float((a - b).abs().max()), 'max rel', float(((a - b).abs() / b.abs().clamp(min=1)).max()))
