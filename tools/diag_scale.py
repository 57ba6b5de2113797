import sys, os, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0,'tests/golden')
from mpntrackseg_b200 import synth, ops
from mpntrackseg_b200.config import default_dataset_params, default_graph_model_params
from mpntrackseg_b200.data.mot_graph import MOTGraph
from mpntrackseg_b200.models.mpn import MOTMPNet
from oracle import graph_ref, mpn_ref
dev = torch.device('cuda:0')
win = synth.make_window(T=15, D=150, k=50, seed=80, node_feats='pooled')
ds = default_dataset_params(top_k_nns=50, frames_per_graph=15)
mp = default_graph_model_params(12, 11)
P = synth.make_params(mp, seed=9, gain=1.25, core_only=True)
for k in ('encoder.node_model.fc_layers.2', 'encoder.edge_model.fc_layers.4'):
    P[k + '.weight'] = P[k + '.weight'] * 32.0; P[k + '.bias'] = P[k + '.bias'] * 32.0
ref_g = graph_ref.build_graph(win.frame, win.reid, synth.det_columns(win), win.fps, ds)
P64 = {k: v.double() for k, v in P.items()}
with torch.no_grad():
    ref = mpn_ref.mpn_forward(P, mp, win.x, ref_g['edge_index'], ref_g['edge_attr'])
    ref64 = mpn_ref.mpn_forward(P64, mp, win.x.double(), ref_g['edge_index'], ref_g['edge_attr'].double())
exp = torch.stack([t.view(-1) for t in ref['classified_edges']]).double().numpy()
exp64 = torch.stack([t.view(-1) for t in ref64['classified_edges']]).numpy()
g = MOTGraph.from_tensors(synth.det_columns(win), win.reid, win.x, None, {'fps': win.fps}, ds).construct_graph_object()
def err(a, b):
    tol = 1e-3 * np.maximum(1.0, np.abs(b)); d = np.abs(a - b)
    return int((d > tol).sum()), float(d.max()), float((d / np.maximum(1, np.abs(b))).max())
print('oracle fp32 vs fp64:', err(exp, exp64))
for eng in ('fp32', 'tc'):
    model = MOTMPNet(mp).to(dev).eval(); model.engine = eng; model.load_state_dict(P, strict=False)
    with torch.no_grad():
        out = model(g)
    got = torch.stack([t.view(-1) for t in out['classified_edges']]).cpu().double().numpy()
    print(eng, 'vs fp32 oracle', err(got, exp), 'vs fp64', err(got, exp64), 'per-step bad vs fp64', [(int((np.abs(got[i]-exp64[i]) > 1e-3*np.maximum(1,np.abs(exp64[i]))).sum())) for i in range(11)])
model = MOTMPNet(mp).to(dev).eval(); model.engine = 'tc'; model.load_state_dict(P, strict=False)
with torch.no_grad():
    out = model(g, return_state=True)
got = torch.stack([t.view(-1) for t in out['classified_edges']]).cpu().double().numpy()
print('tc return_state vs fp32 oracle', err(got, exp))
with torch.no_grad():
    out2 = model(g)
got2 = torch.stack([t.view(-1) for t in out2['classified_edges']]).cpu().double().numpy()
print('again without state: equal to first tc?', err(got2, exp), float(np.abs(got2-got).max()))
