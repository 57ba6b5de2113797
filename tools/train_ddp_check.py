"""Multi-GPU check of the training step (run under torchrun, one rank per GPU):
every rank trains on its own KITTI-shaped window; gradients are summed with ONE NCCL all-reduce of the flat
bucket and averaged inside the Adam kernel.  Rank 0 replays the same windows in one process
(accumulate -> mean -> Adam) and compares the parameters.  Prints the measured step time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from mpntrackseg_b200 import synth
from mpntrackseg_b200.config import default_dataset_params, default_graph_model_params
from mpntrackseg_b200.data.mot_graph import MOTGraph
from mpntrackseg_b200.models.mpn import MOTMPNet
from mpntrackseg_b200.training import CoreTrainer

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
ds = default_dataset_params(top_k_nns=100, frames_per_graph=20)
mp = default_graph_model_params(12, 11)
P = synth.make_params(mp, seed=6, gain=0.95, core_only=True)


def window(seed):
    w = synth.make_window(T=20, D=8, k=100, seed=seed)
    g = MOTGraph.from_tensors(synth.det_columns(w), w.reid, w.x.to(dev), None, {'fps': 30.0}, ds).construct_graph_object()
    ident = w.ident.to(dev)
    return g, (ident[g.edge_index[0]] == ident[g.edge_index[1]]).float()


def fresh():
    m = MOTMPNet(mp).to(dev)
    m.load_state_dict(P, strict=False)
    return CoreTrainer(m)

tr = fresh()
g, labels = window(100 + rank)
# first step by hand to keep the all-reduced gradient bucket for the comparison below
tr.loss_and_grads(g, labels)
tr.all_reduce_grads()
grad_ddp = tr.grad.clone()
tr.adam_step(grad_scale=1.0 / world)
for it in range(2):
    loss = tr.train_step(g, labels)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for it in range(10):
    loss = tr.train_step(g, labels)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
if rank == 0:
    ref = fresh()
    graphs = [window(100 + r) for r in range(world)]
    gerr = None
    for it in range(13):
        ref.grad.zero_()
        for gg, ll in graphs:
            ref.loss_and_grads(gg, ll, zero_grad=False)
        if gerr is None:
            gerr = float((ref.grad - grad_ddp).abs().max()) / float(ref.grad.abs().max())
        ref.adam_step(grad_scale=1.0 / world)
    err = float((ref.flat - tr.flat).abs().max()) / float(ref.flat.abs().max())
    e = g.edge_index.shape[1]
    print(f'world={world} E={e} train step {dt * 1e3:.2f} ms  ({12 * e * world / dt / 1e6:.1f} M edge-updates/s fwd+bwd)  '
          f'loss={float(loss):.5f}  all-reduced gradient vs single-process sum: {gerr:.2e}; params after 13 Adam steps: {err:.2e}')
    # gradients agree to fp32 summation order; Adam amplifies noise-level gradients of dead units (update = +-lr),
    # so the parameter comparison is loose
    assert gerr < 1e-5 and err < 5e-3, (gerr, err)
if world > 1:
    dist.destroy_process_group()
