"""Development: per-phase cycle trace of the tensor-core edge kernel (CTA 0, group 0)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpntrackseg_b200 import synth, _cabi
from mpntrackseg_b200.config import default_dataset_params, default_graph_model_params
from mpntrackseg_b200.data.mot_graph import MOTGraph
from mpntrackseg_b200.models.mpn import MOTMPNet
dev = torch.device('cuda')
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ds = default_dataset_params(50, 15); mp = default_graph_model_params(12, 11)
P = synth.make_params(mp, seed=9, gain=1.25, core_only=True)
model = MOTMPNet(mp).to(dev).eval(); model.load_state_dict(P, strict=False); model.engine = 'tc'
graphs = []
for g in range(G):
    w = synth.make_window(T=15, D=150, k=50, seed=g)
    graphs.append(MOTGraph.from_tensors(synth.det_columns(w), w.reid, w.x.to(dev), None, {'fps': 30.0}, ds).construct_graph_object())
buf = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
lib = _cabi.lib()
lib.mpn_tc_set_trace.argtypes = [C.c_void_p]
with torch.no_grad():
    model.forward_batch(graphs)
    lib.mpn_tc_set_trace(C.c_void_p(buf.data_ptr()))
    model.forward_batch(graphs)
    lib.mpn_tc_set_trace(None)
torch.cuda.synchronize()
if os.environ.get('MPN_TC_VARIANT', '3') != '2':
    sys.exit('the cycle trace is instrumented in the previous kernel only: run with MPN_TC_VARIANT=2')
t = buf.cpu().view(64, 16)
names = ['start', 'cpasync_wait', 'load->arrive', 'prow issued', 'dready1', 'epi1 done', 'dready2', 'epi2 arrive', 'cls done', 'dready3', 'epi3 done', 'dready4', 'tile end']
for i in range(2, 12):
    row = t[i]
    if row[0] == 0: break
    d = [int(row[k] - row[k - 1]) for k in range(1, 13)]
    print(f'tile {i:2d} total {int(row[12]-row[0]):6d} | ' + ' '.join(f'{n}:{v}' for n, v in zip(names[1:], d)))
