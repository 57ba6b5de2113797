"""Model / dataset hyper-parameter dictionaries and the parameter-shape contract.

The dictionaries have the same keys as the reference's ``graph_model_params`` and
``dataset_params`` (reference: configs/tracking_cfg.yaml:64-85,134-218), so a
``hparams['graph_model_params']`` dict loaded from the reference's YAML can be passed
to :class:`mpntrackseg_b200.models.mpn.MOTMPNet` unchanged.

``param_shapes`` states, key by key, the ``state_dict`` layout that the reference
model exposes (reference: models/mpn.py:220-331, models/mlp.py:4-28, models/cnn.py:4-84).
It is the contract that lets ``mots20.ckpt`` / ``kitti.ckpt`` load into the CUDA model.
"""
from collections import OrderedDict
from copy import deepcopy

EDGE_FEATS = ('secs_time_dists', 'norm_feet_x_dists', 'norm_feet_y_dists',
              'bb_height_dists', 'bb_width_dists', 'emb_dist')


def default_graph_model_params(num_enc_steps=12, num_class_steps=11):
    """Shipped widths (reference: configs/tracking_cfg.yaml:134-218) with the step
    counts BASELINE.json quotes its metric on (12 message-passing steps)."""
    return {
        'node_agg_fn': 'sum',
        'num_enc_steps': num_enc_steps,
        'num_class_steps': num_class_steps,
        'reattach_initial_nodes': True,
        'reattach_initial_edges': True,
        'encoder_feats_dict': {
            'edge_in_dim': 6, 'edge_dims': [18, 18], 'edge_out_dim': 16,
            'node_in_dim': 2048, 'node_dims': [128], 'node_out_dim': 32,
            'dropout_p': 0, 'use_batchnorm': False},
        'edge_model_feats_dict': {'dims': [80, 16], 'dropout_p': 0, 'use_batchnorm': False},
        'node_model_feats_dict': {'dims': [56, 32], 'dropout_p': 0, 'use_batchnorm': False},
        'classifier_feats_dict': {
            'edge_in_dim': 16, 'edge_dims': [8], 'edge_out_dim': 1,
            'dropout_p': 0, 'use_batchnorm': False},
        'node_ext_encoder_feats_dict': {
            'input_dim': 256, 'dims': [128, 32], 'kernel_sizes': [1, 1], 'strides': [1, 1],
            'paddings': [0, 0], 'dropout_p': 0, 'use_batchnorm': False},
        'attention_model_feats_dict': {'fc_dims': [16, 1], 'dropout_p': 0, 'use_batchnorm': False},
        'node_ext_model_feats_dict': {
            'dims': [96, 32], 'kernel_sizes': [3, 3], 'strides': [1, 1], 'paddings': [1, 1],
            'dropout_p': 0, 'use_batchnorm': False},
        'mask_model_feats_dict': {
            'feature_encoder_feats_dict': {
                'input_dim': 256, 'dims': [32], 'kernel_sizes': [1], 'strides': [1],
                'paddings': [0], 'dropout_p': 0, 'use_batchnorm': False},
            'mask_head_feats_dict': {
                'input_dim': 64, 'dims': [64, 64, 64], 'kernel_sizes': [3, 3, 3],
                'strides': [1, 1, 1], 'paddings': [1, 1, 1], 'dropout_p': 0,
                'use_batchnorm': False},
            'mask_predictor_feats_dict': {
                'input_dim': 64, 'dims': [64, 64, 64, 1], 'kernel_sizes': [2, 3, 2, 1],
                'strides': [2, 1, 2, 1], 'paddings': [0, 1, 0, 0],
                'transposed': [True, False, True, False]}},
    }


def default_dataset_params(top_k_nns=50, frames_per_graph=15, reciprocal_k_nns=True):
    """Graph-construction subset of ``dataset_params``
    (reference: configs/tracking_cfg.yaml:64-85)."""
    return {
        'frames_per_graph': frames_per_graph,
        'max_frame_dist': 'max',
        'max_detects': None,
        'top_k_nns': top_k_nns,
        'reciprocal_k_nns': reciprocal_k_nns,
        'edge_feats_to_use': list(EDGE_FEATS),
    }


def _mlp_shapes(prefix, in_dim, dims, out, use_batchnorm=False, dropout_p=0):
    """Sequential slots of ``MLP``: Linear [, BatchNorm1d], ReLU [, Dropout] per hidden layer; a layer of width 1 is
    a bare Linear (reference: models/mlp.py:12-23)."""
    slot = 0
    for d in dims:
        out[f'{prefix}.fc_layers.{slot}.weight'] = (d, in_dim)
        out[f'{prefix}.fc_layers.{slot}.bias'] = (d,)
        slot += 1
        if d != 1:
            if use_batchnorm:
                for name, shape in (('weight', (d,)), ('bias', (d,)), ('running_mean', (d,)), ('running_var', (d,)),
                                    ('num_batches_tracked', ())):
                    out[f'{prefix}.fc_layers.{slot}.{name}'] = shape
                slot += 1
            slot += 1                                       # ReLU
            if dropout_p != 0:
                slot += 1
        in_dim = d


def _cnn_shapes(prefix, in_dim, dims, ks, out, transposed=None):
    """Conv layers sit at even Sequential slots (reference: models/cnn.py:25-41,70-82)."""
    for i, (d, k) in enumerate(zip(dims, ks)):
        if transposed is not None and transposed[i]:
            out[f'{prefix}.layers.{2 * i}.weight'] = (in_dim, d, k, k)
        else:
            out[f'{prefix}.layers.{2 * i}.weight'] = (d, in_dim, k, k)
        out[f'{prefix}.layers.{2 * i}.bias'] = (d,)
        in_dim = d


def core_dims(model_params):
    """Widths of the core message-passing path, derived as the reference derives them
    (reference: models/mpn.py:275-287)."""
    enc = model_params['encoder_feats_dict']
    nf = 2 if model_params['reattach_initial_nodes'] else 1
    ef = 2 if model_params['reattach_initial_edges'] else 1
    dn, de = enc['node_out_dim'], enc['edge_out_dim']
    return {
        'node_factor': nf, 'edge_factor': ef, 'dn': dn, 'de': de,
        'edge_mlp_in': nf * 2 * dn + ef * de,
        'flow_mlp_in': nf * dn + de,
        'edge_mlp_dims': list(model_params['edge_model_feats_dict']['dims']),
        'flow_mlp_dims': list(model_params['node_model_feats_dict']['dims']),
        'enc_edge_dims': [enc['edge_in_dim']] + list(enc['edge_dims']) + [enc['edge_out_dim']],
        'enc_node_dims': [enc['node_in_dim']] + list(enc['node_dims']) + [enc['node_out_dim']],
    }


def param_shapes(model_params, core_only=False):
    """``OrderedDict`` name -> shape in the reference's ``state_dict`` order."""
    p = model_params
    enc = p['encoder_feats_dict']
    cls = p['classifier_feats_dict']
    d = core_dims(p)
    out = OrderedDict()
    bn = lambda dct: dict(use_batchnorm=dct.get('use_batchnorm', False), dropout_p=dct.get('dropout_p', 0))
    _mlp_shapes('encoder.node_model', enc['node_in_dim'],
                list(enc['node_dims']) + [enc['node_out_dim']], out, **bn(enc))
    _mlp_shapes('encoder.edge_model', enc['edge_in_dim'],
                list(enc['edge_dims']) + [enc['edge_out_dim']], out, **bn(enc))
    _mlp_shapes('classifier.edge_model', cls['edge_in_dim'],
                list(cls['edge_dims']) + [cls['edge_out_dim']], out, **bn(cls))
    if not core_only:
        ne = p['node_ext_encoder_feats_dict']
        _cnn_shapes('node_ext_encoder', ne['input_dim'], ne['dims'], ne['kernel_sizes'], out)
        mm = p['mask_model_feats_dict']
        fe, mh, mp = (mm['feature_encoder_feats_dict'], mm['mask_head_feats_dict'],
                      mm['mask_predictor_feats_dict'])
        _cnn_shapes('mask_predictor.feature_encoder', fe['input_dim'], fe['dims'],
                    fe['kernel_sizes'], out)
        out['mask_predictor.layer_norm.weight'] = (64, 14, 14)
        out['mask_predictor.layer_norm.bias'] = (64, 14, 14)
        _cnn_shapes('mask_predictor.mask_head', mh['input_dim'], mh['dims'],
                    mh['kernel_sizes'], out)
        _cnn_shapes('mask_predictor.mask_predictor', mp['input_dim'], mp['dims'],
                    mp['kernel_sizes'], out, transposed=mp['transposed'])
    _mlp_shapes('MPNet.edge_model.edge_model', d['edge_mlp_in'], d['edge_mlp_dims'], out, **bn(p['edge_model_feats_dict']))
    _mlp_shapes('MPNet.node_model.flow_in_model', d['flow_mlp_in'], d['flow_mlp_dims'], out, **bn(p['node_model_feats_dict']))
    _mlp_shapes('MPNet.node_model.flow_out_model', d['flow_mlp_in'], d['flow_mlp_dims'], out, **bn(p['node_model_feats_dict']))
    out['MPNet.node_model.node_model.0.weight'] = (d['dn'], 2 * d['dn'])
    out['MPNet.node_model.node_model.0.bias'] = (d['dn'],)
    if not core_only:
        nx = p['node_ext_model_feats_dict']
        in_dim = 3 * p['node_ext_encoder_feats_dict']['dims'][-1] * d['node_factor']
        _cnn_shapes('MPAttentionNet.node_model', in_dim, nx['dims'], nx['kernel_sizes'], out)
    return out


def clone_params(model_params):
    return deepcopy(model_params)
