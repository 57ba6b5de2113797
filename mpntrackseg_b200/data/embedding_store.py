"""Per-sequence store of precomputed appearance embeddings and its loader.

reference: src/mot_neural_solver/data/seq_processing/seq_processor.py:440-475 (ReID ``[n,1+256]`` and node-core
``[n,1+2048,8,4]`` tensors saved per frame as ``<frame>.pt``, detection id in column / channel 0), :545-557 (node-ext
``[n,1+256,14,14]``), and src/mot_neural_solver/utils/rgb.py:150-188 ``load_precomputed_embeddings`` (load the frames of
a window, drop the rows whose detection id is not in the window's table, check the order, strip the id).

Two layers:

* ``EmbeddingStore`` writes / reads the reference's on-disk layout unchanged (so a store produced by the reference's
  preprocessing is read as is), and ``load_precomputed_embeddings`` keeps the reference's signature and result.
* The reference stores the node-core embedding as a 2048x8x4 map per detection (262 KB) and ``MOTMPNet.forward``
  average-pools it first thing (models/mpn.py:351-352).  ``EmbeddingStore.pool`` writes the pooled ``[n,1+2048]``
  variant once (32x smaller, pooled by the GPU kernel that the forward would run, so the values are the ones the
  forward computes), and ``SequenceEmbeddings`` packs a sequence's rows into page-locked host memory in table order:
  a frame window is then a contiguous slice that goes host -> device in one asynchronous copy.

Host-side IO and bookkeeping only; the arithmetic (pooling) is the library's kernel.
"""
import os
import os.path as osp

import numpy as np
import torch

from .. import ops


def _col(table, name):
    v = table[name]
    if torch.is_tensor(v):
        return v.cpu().numpy()
    return np.asarray(v.values if hasattr(v, 'values') else v)


class EmbeddingStore(object):
    """``<seq_path>/processed_data/embeddings/<det_file_name>/<embeddings_dir>/<frame>.pt``
    (seq_processor.py:448-451, rgb.py:166-168)."""

    def __init__(self, seq_info_dict):
        self.seq_info_dict = seq_info_dict
        self.root = osp.join(seq_info_dict['seq_path'], 'processed_data', 'embeddings', seq_info_dict['det_file_name'])

    def dir(self, embeddings_dir):
        return osp.join(self.root, embeddings_dir)

    def frames(self, embeddings_dir):
        return sorted(int(f[:-3]) for f in os.listdir(self.dir(embeddings_dir)) if f.endswith('.pt'))

    def write(self, embeddings_dir, frame_nums, det_ids, embeds):
        """Save ``embeds`` ([n,C] or [n,C,H,W], any order) grouped by frame with the detection id prepended as column 0
        (1-D) or as a constant channel 0 (3-D).  seq_processor.py:445-446,462-472 / :545-557"""
        frame_nums = torch.as_tensor(np.asarray(frame_nums)).view(-1)
        det_ids = torch.as_tensor(np.asarray(det_ids)).view(-1)
        embeds = embeds.detach().cpu().float()
        if embeds.dim() == 2:
            tagged = torch.cat((det_ids.view(-1, 1).float(), embeds), dim=1)
        elif embeds.dim() == 4:
            tag = det_ids.view(-1, 1, 1, 1).float().expand(-1, -1, embeds.shape[2], embeds.shape[3])
            tagged = torch.cat((tag, embeds), dim=1)
        else:
            raise ValueError(f'embeddings must be [n,C] or [n,C,H,W], got {tuple(embeds.shape)}')
        path = self.dir(embeddings_dir)
        os.makedirs(path, exist_ok=True)
        for frame in torch.unique(frame_nums).tolist():
            torch.save(tagged[frame_nums == frame].clone(), osp.join(path, f'{int(frame)}.pt'))

    def read_frame(self, embeddings_dir, frame):
        return torch.load(osp.join(self.dir(embeddings_dir), f'{int(frame)}.pt'))

    def pool(self, src_dir, dst_dir, device=None):
        """Write the pooled variant of a 3-D store: every ``[n,1+C,H,W]`` frame tensor becomes ``[n,1+C]`` with
        ``AdaptiveAvgPool2d((1,1))`` applied by the library's kernel (models/mpn.py:351-352).  One pass, offline."""
        dev = device or torch.device('cuda')
        os.makedirs(self.dir(dst_dir), exist_ok=True)
        for frame in self.frames(src_dir):
            t = self.read_frame(src_dir, frame)
            if t.dim() != 4:
                raise ValueError(f'{src_dir}/{frame}.pt is not a 3-D embedding tensor')
            pooled = ops.avgpool(t[:, 1:].contiguous().to(dev)).cpu() if t.shape[0] else t.new_zeros((0, t.shape[1] - 1))
            torch.save(torch.cat((t[:, :1, 0, 0], pooled), dim=1), osp.join(self.dir(dst_dir), f'{frame}.pt'))
        return dst_dir


def load_precomputed_embeddings(det_df, seq_info_dict, embeddings_dir, use_cuda, embedding_dim='1D', pin_memory=False):
    """Drop-in for utils/rgb.py:150-188: embeddings of the detections in ``det_df`` (rows sorted by frame, detection id),
    read from the per-frame files; ``embeddings_dir`` is relative to ``<seq_path>/processed_data`` like the reference's.
    ``pin_memory`` (extension): build the result in page-locked memory so that the device copy is asynchronous."""
    assert embedding_dim in ['1D', '3D'], "Embedding dimension is not valid!"
    embeddings_path = osp.join(seq_info_dict['seq_path'], 'processed_data', embeddings_dir)
    frames_to_retrieve = sorted(np.unique(_col(det_df, 'frame')).tolist())
    parts = [torch.load(osp.join(embeddings_path, f'{int(f)}.pt')) for f in frames_to_retrieve]
    want = _col(det_df, 'detection_id').astype(np.int64)
    ids = torch.cat([(p[:, 0] if p.dim() == 2 else p[:, 0, 0, 0]) for p in parts]).to(torch.int64).numpy()
    keep = np.isin(ids, want)
    assert_str = ("Problems loading embeddings. Indices between query and stored embeddings do not match. "
                  "BOTH SHOULD BE SORTED!")
    assert keep.sum() == want.shape[0] and (ids[keep] == want).all(), assert_str
    for p in parts:
        if (p.dim() == 2) != (embedding_dim == '1D'):
            raise ValueError(f"stored tensors are {p.dim() - 1}-D, embedding_dim={embedding_dim!r}")
    shape = (int(keep.sum()),) + (tuple(parts[0].shape[1:]) if parts else (0,))
    shape = (shape[0], shape[1] - 1) + shape[2:]
    out = torch.empty(shape, dtype=torch.float32, pin_memory=bool(pin_memory))
    off, kofs = 0, 0
    for p in parts:                                    # strip the id column / channel while copying, frame by frame
        k = torch.from_numpy(keep[kofs:kofs + p.shape[0]])
        kofs += p.shape[0]
        sel = p[k][:, 1:]
        out[off:off + sel.shape[0]].copy_(sel)
        off += sel.shape[0]
    if use_cuda and torch.cuda.is_available():
        return out.to(torch.device('cuda'), non_blocking=bool(pin_memory))
    return out


class SequenceEmbeddings(object):
    """A sequence's appearance data packed in page-locked host memory, rows in detection-table order
    (sorted by frame, detection id): ``reid [N,256]``, ``node_core [N,2048]`` (pooled store) or ``[N,2048,8,4]``,
    optionally ``node_ext [N,256,14,14]``; ``frame_ptr`` maps a frame number to its row range, so the window
    ``[start_frame, end_frame]`` is the contiguous slice ``rows(start_frame, end_frame)``."""

    def __init__(self, det_df, seq_info_dict, dataset_params, pooled=True, with_ext=False, pin_memory=True):
        emb_dir = osp.join('embeddings', seq_info_dict['det_file_name'])
        core_dir = dataset_params['node_core_embeddings_dir'] + ('_pooled' if pooled else '')
        load = lambda d, dim: load_precomputed_embeddings(det_df, seq_info_dict, osp.join(emb_dir, d), use_cuda=False,
                                                          embedding_dim=dim, pin_memory=pin_memory)
        self.reid = load(dataset_params['reid_embeddings_dir'], '1D')
        self.node_core = load(core_dir, '1D' if pooled else '3D')
        self.node_ext = load(dataset_params['node_ext_embeddings_dir'], '3D') if with_ext else None
        frames = _col(det_df, 'frame').astype(np.int64)
        assert (np.diff(frames) >= 0).all(), 'the detection table must be sorted by frame'
        self.frames = np.unique(frames)
        self.frame_ptr = np.concatenate((np.searchsorted(frames, self.frames, side='left'), [frames.shape[0]]))

    def rows(self, start_frame, end_frame):
        a = int(np.searchsorted(self.frames, start_frame, side='left'))
        b = int(np.searchsorted(self.frames, end_frame, side='right'))
        return int(self.frame_ptr[a]), int(self.frame_ptr[b])

    def window(self, start_frame, end_frame, device=None, stream=None):
        """(reid, node_core[, node_ext]) of the window on ``device``: one asynchronous copy per tensor from the pinned
        slices, enqueued on ``stream`` (default: the current stream)."""
        dev = device or torch.device('cuda')
        a, b = self.rows(start_frame, end_frame)
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            out = [self.reid[a:b].to(dev, non_blocking=True), self.node_core[a:b].to(dev, non_blocking=True)]
            if self.node_ext is not None:
                out.append(self.node_ext[a:b].to(dev, non_blocking=True))
        return tuple(out)


class _null(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
