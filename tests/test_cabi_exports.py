"""CPU-side checks: the shared library loads and exports every symbol the header declares,
the ctypes binding covers all of them, and the parameter-shape contract matches the module."""
import ctypes
import os
import re

from mpntrackseg_b200 import _cabi
from mpntrackseg_b200.config import default_graph_model_params, param_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mpntrack_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mpn_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    syms = _declared_symbols()
    assert len(syms) >= 18
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f'{s} declared in include/mpntrack_b200.h but not exported'


def test_binding_covers_header():
    assert sorted(_cabi.SIGNATURES) == _declared_symbols()
    assert _cabi.lib().mpn_abi_version() == _cabi.ABI_VERSION


def test_argument_errors_are_reported_without_a_gpu():
    lib = _cabi.lib()
    # null descriptors -> MPN_EINVAL with a message, no CUDA call involved
    rc = lib.mpn_mp_forward(None, None, None, None, 1, 1, None, None, None, None, None)
    assert rc == -1 and b'null descriptor' in lib.mpn_last_error()
    rc = lib.mpn_edge_feats_assemble(None, None, 4, None, None, None, None, None, 30.0, None, 6, None, None, None)
    assert rc == -1


def test_state_dict_contract():
    import torch  # noqa: F401
    from mpntrackseg_b200.models.mpn import MOTMPNet
    mp = default_graph_model_params()
    sd = MOTMPNet(mp).state_dict()
    full = param_shapes(mp)
    assert list(sd) == list(full)                     # same keys in the reference's registration order
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(full[k]), k
    core = param_shapes(mp, core_only=True)
    assert sum(int(torch.tensor(s).prod()) for s in full.values()) == 740966
    assert sum(int(torch.tensor(s).prod()) for s in core.values()) == 296293


def test_cpu_tensors_are_rejected():
    import pytest
    import torch
    from mpntrackseg_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.linear(torch.zeros(2, 3), torch.zeros(4, 3), None, True)
