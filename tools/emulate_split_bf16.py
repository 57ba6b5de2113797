"""Numerics probe (development): emulate the planned tensor-core kernel's arithmetic on CPU --
split-bf16 operands (hi/lo, 3 products, fp32 accumulate), x[row] part of the edge MLP hoisted
in fp32 -- and compare edge logits with the fp32 oracle on golden / dense cases."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden'))
import numpy as np, torch
import torch.nn.functional as F
from mpntrackseg_b200 import synth
from mpntrackseg_b200.config import default_dataset_params, default_graph_model_params
from oracle import graph_ref, mpn_ref

def split(t, dt=torch.bfloat16):
    hi = t.to(dt).float()
    lo = (t - hi).to(dt).float()
    return hi, lo

def lin_split(x, w, b=None, mode='split3'):
    if mode == 'fp32':
        return F.linear(x, w, b)
    dt = torch.float16 if mode.endswith('fp16') else torch.bfloat16
    xh, xl = split(x, dt); wh, wl = split(w, dt)
    if mode == 'bf16':
        y = xh @ wh.T
    elif mode.startswith('split2w'):                               # weights rounded once, activations split
        y = xh @ wh.T + xl @ wh.T
    else:
        y = xh @ wh.T + xh @ wl.T + xl @ wh.T
    return y if b is None else y + b

def forward(P, mp, x, ei, ea, mode):
    x0, e0 = mpn_ref.encode(P, x, ea)
    row, col = ei
    out_m, in_m = row < col, row > col
    W0, b0 = P['MPNet.edge_model.edge_model.fc_layers.0.weight'], P['MPNet.edge_model.edge_model.fc_layers.0.bias']
    W1, b1 = P['MPNet.edge_model.edge_model.fc_layers.2.weight'], P['MPNet.edge_model.edge_model.fc_layers.2.bias']
    fl = {d: [P[f'MPNet.node_model.flow_{d}_model.fc_layers.{s}.{t}'] for s in (0, 2) for t in ('weight', 'bias')] for d in ('in', 'out')}
    Wn, bn = P['MPNet.node_model.node_model.0.weight'], P['MPNet.node_model.node_model.0.bias']
    xs, es = x0, e0
    n = x0.shape[0]
    logits = []
    for step in range(mp['num_enc_steps']):
        xc = torch.cat((x0, xs), 1)
        prow = F.linear(xc, W0[:, :64], b0)                       # hoisted, exact fp32
        a = torch.cat((xc[col], e0, es), 1)                        # K = 96 on tensor cores
        h = F.relu(prow[row] + lin_split(a, W0[:, 64:], None, mode))
        e2 = F.relu(lin_split(h, W1, b1, mode))
        flows = {}
        for d, m in (('out', out_m), ('in', in_m)):
            w0, bb0, w1, bb1 = fl[d]
            g = F.relu(lin_split(torch.cat((xc[col[m]], e2[m]), 1), w0, bb0, mode))
            msg = F.relu(lin_split(g, w1, bb1, mode))
            flows[d] = mpn_ref.segment_add(msg, row[m], n)
        xs = F.relu(F.linear(torch.cat((flows['in'], flows['out']), 1), Wn, bn))
        es = e2
        logits.append(mpn_ref.classify(P, es))
    return logits[-1].view(-1)

def run(T, D, k, gain, seed=0, wseed=9):
    win = synth.make_window(T=T, D=D, k=k, seed=seed)
    ds = default_dataset_params(k, T); mp = default_graph_model_params(12, 11)
    g = graph_ref.build_graph(win.frame, win.reid, synth.det_columns(win), win.fps, ds)
    P = synth.make_params(mp, seed=wseed, gain=gain, core_only=True)
    with torch.no_grad():
        ref = mpn_ref.mpn_forward(P, mp, win.x, g['edge_index'], g['edge_attr'])['classified_edges'][-1].view(-1)
        shift = ref.median()
        for mode in ('fp32', 'split3_fp16', 'split2w_fp16'):
            got = forward(P, mp, win.x, g['edge_index'], g['edge_attr'], mode)
            err = (got - ref).abs()
            rel = (err / ref.abs().clamp(min=1)).max()
            pr, pg = torch.sigmoid(ref - shift), torch.sigmoid(got - shift)
            dec = ((pr > .5) != (pg > .5)) & ((pr - .5).abs() > 1e-3)
            print(f'T={T} D={D} k={k} gain={gain} E={g["edge_index"].shape[1]} logit_std={ref.std():.3g} mode={mode}: '
                  f'max|dlogit|={err.max():.3e} max rel={rel:.3e} max|dp|={(pr-pg).abs().max():.3e} flips={int(dec.sum())}')

if __name__ == '__main__':
    run(15, 30, 50, 1.2)
    run(20, 8, 100, 0.95)
    run(15, 150, 50, 1.25)
