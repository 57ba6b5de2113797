"""Development check: AttnAggregate backward kernels against torch autograd of the same aggregation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpntrackseg_b200 import ops
from mpntrackseg_b200.training import AttnAggregate
from oracle import mpn_ref

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
n, F = 23, 7 * 5
ii, jj = torch.triu_indices(n, n, offset=1)
keep = torch.rand(ii.numel(), generator=g) < 0.3
pi, pj = ii[keep], jj[keep]
ei = torch.cat((torch.stack((pi, pj)), torch.stack((pj, pi))), dim=1)
perm = torch.randperm(ei.shape[1], generator=g)
ei = ei[:, perm].contiguous()
z = torch.randn(n, 7, 5, 1, generator=g)
lg = torch.randn(ei.shape[1], generator=g)
G_in, G_out = torch.randn(n, 7, 5, 1, generator=g), torch.randn(n, 7, 5, 1, generator=g)

def ref(z, lg):
    src, dst = ei[0], ei[1]
    flows = {}
    for name, sel in (('out', src < dst), ('in', src > dst)):
        w = mpn_ref.segment_softmax(lg[sel].view(-1, 1), src[sel])
        flows[name] = mpn_ref.segment_add(z[dst[sel]] * w[:, :, None, None], src[sel], n)
    return flows['in'], flows['out']

zr, lr = z.clone().requires_grad_(True), lg.clone().requires_grad_(True)
fi, fo = ref(zr, lr)
((fi * G_in).sum() + (fo * G_out).sum()).backward()

lay = ops.edge_layout(ei.to(dev), n)
e = lay.num_edges
scol = lay.slot_col[:e]
perm_c = torch.argsort(scol, stable=True).to(torch.int32)
ptr_c = torch.zeros(n + 1, dtype=torch.int32, device=dev)
ptr_c[1:] = torch.cumsum(torch.bincount(scol.long(), minlength=n), 0).to(torch.int32)
zc, lc = z.to(dev).requires_grad_(True), lg.to(dev).requires_grad_(True)
ci, co = AttnAggregate.apply(zc, lc, lay, perm_c, ptr_c)
print('fwd err', float((ci.cpu() - fi).abs().max()), float((co.cpu() - fo).abs().max()))
((ci * G_in.to(dev)).sum() + (co * G_out.to(dev)).sum()).backward()
print('dz err', float((zc.grad.cpu() - zr.grad).abs().max()), 'scale', float(zr.grad.abs().max()))
print('dl err', float((lc.grad.cpu() - lr.grad).abs().max()), 'scale', float(lr.grad.abs().max()))
bad = (zc.grad.cpu() - zr.grad).abs().flatten(1).max(dim=1).values
print('nodes with dz error > 1e-4:', torch.nonzero(bad > 1e-4).view(-1).tolist())
