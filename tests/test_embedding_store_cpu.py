"""Embedding store against a fixture made by the reference's own ``load_precomputed_embeddings``
(utils/rgb.py:150-188) on files in the layout its preprocessing writes (seq_processor.py:445-446,462-472).
Host-side IO only: no GPU needed (the pooling pass is covered by the GPU tests)."""
import os
import os.path as osp

import numpy as np
import pytest
import torch

from mpntrackseg_b200.data.embedding_store import EmbeddingStore, load_precomputed_embeddings

GOLD = dict(np.load(osp.join(osp.dirname(__file__), 'golden', 'embedding_store.npz')))


def _write_store(tmp_path):
    seq_info = {'seq_path': str(tmp_path), 'det_file_name': 'det'}
    store = EmbeddingStore(seq_info)
    store.write('reid', GOLD['frames'], GOLD['det_ids'], torch.from_numpy(GOLD['reid']))
    store.write('core', GOLD['frames'], GOLD['det_ids'], torch.from_numpy(GOLD['core']))
    return seq_info, store


def test_store_files_have_the_reference_layout(tmp_path):
    _, store = _write_store(tmp_path)
    assert store.frames('reid') == [4, 5, 7] and store.frames('core') == [4, 5, 7]
    assert np.array_equal(store.read_frame('reid', 5).numpy(), GOLD['file_reid_5'])       # [n, 1+256]: id in column 0
    assert np.array_equal(store.read_frame('core', 7).numpy(), GOLD['file_core_7'])       # [n, 1+C, H, W]: id in channel 0
    assert osp.isfile(osp.join(str(tmp_path), 'processed_data', 'embeddings', 'det', 'reid', '4.pt'))


@pytest.mark.parametrize('pin', [False, True])
def test_loader_matches_the_reference_loader(tmp_path, pin):
    seq_info, _ = _write_store(tmp_path)
    keep = GOLD['keep']
    df = {'frame': GOLD['frames'][keep], 'detection_id': GOLD['det_ids'][keep]}
    pin = pin and torch.cuda.is_available()                  # page-locking needs a CUDA runtime
    reid = load_precomputed_embeddings(df, seq_info, osp.join('embeddings', 'det', 'reid'), use_cuda=False, pin_memory=pin)
    core = load_precomputed_embeddings(df, seq_info, osp.join('embeddings', 'det', 'core'), use_cuda=False,
                                       embedding_dim='3D', pin_memory=pin)
    assert np.array_equal(reid.numpy(), GOLD['out_reid']) and np.array_equal(core.numpy(), GOLD['out_core'])
    sub = {k: v[df['frame'] != 5] for k, v in df.items()}    # a window that skips a stored frame
    out = load_precomputed_embeddings(sub, seq_info, osp.join('embeddings', 'det', 'reid'), use_cuda=False)
    assert np.array_equal(out.numpy(), GOLD['out_sub'])


def test_loader_rejects_an_unsorted_or_unknown_table(tmp_path):
    seq_info, _ = _write_store(tmp_path)
    bad = {'frame': np.array([4, 4]), 'detection_id': np.array([12, 10])}                # not sorted by detection id
    with pytest.raises(AssertionError):
        load_precomputed_embeddings(bad, seq_info, osp.join('embeddings', 'det', 'reid'), use_cuda=False)
    missing = {'frame': np.array([4]), 'detection_id': np.array([99])}
    with pytest.raises(AssertionError):
        load_precomputed_embeddings(missing, seq_info, osp.join('embeddings', 'det', 'reid'), use_cuda=False)


def test_motgraph_reference_constructor_selects_the_window_rows():
    """MOTGraph(seq_det_df, start_frame, end_frame, ...) -> graph_df / frames as data/mot_graph.py:108-147."""
    import pandas as pd
    from mpntrackseg_b200.data.mot_graph import MOTGraph
    frames = np.repeat(np.arange(1, 11), 3)
    df = pd.DataFrame({'frame': frames, 'detection_id': np.arange(30)[::-1].copy(), 'bb_left': np.zeros(30)})
    dp = {'frames_per_graph': 4, 'max_detects': None, 'max_frame_dist': 'max'}
    g = MOTGraph(seq_det_df=df, start_frame=2, end_frame=8, step_size=3, ensure_end_is_in=True, dataset_params=dp)
    assert g.frames == [2, 5, 8] and list(g.graph_df.frame.unique()) == [2, 5, 8]
    assert (np.diff(g.graph_df.detection_id.values.reshape(3, 3), axis=1) > 0).all()     # sorted inside a frame
    g = MOTGraph(seq_det_df=df, start_frame=3, step_size=2, dataset_params=dp)           # open end: frames_per_graph caps
    assert g.frames == [3, 5, 7, 9]
    g = MOTGraph(seq_det_df=df, start_frame=1, step_size=1, dataset_params=dict(dp, frames_per_graph='max', max_detects=7))
    assert g.frames == [1, 2]                                                            # 3 + 3 <= 7 < 9 detections
