"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/ttk.h declares, and its
host-side logic (network plan, parameter lists, argument validation) agrees with the oracle's description
of the reference.  No kernel is launched here."""
import ctypes as C
import os
import re

import pytest

from oracle import hrnet as ohr
from oracle import uplift as oup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def L():
    from upliftingtabletennis_b200 import _lib
    return _lib


def test_exports_match_header(L):
    header = open(os.path.join(ROOT, 'include', 'ttk.h')).read()
    declared = sorted(set(re.findall(r'TTK_API\s+[\w\s\*]+?\b(ttk_\w+)\s*\(', header)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L.lib, name), 'libttk.so does not export %s' % name
    assert set(declared) == set(L.EXPORTS), 'ctypes binding and header disagree'
    assert L.lib.ttk_version() == 100


@pytest.mark.parametrize('in_ch,out_ch', [(9, 3), (3, 13)])
def test_hrnet_plan_matches_reference_layout(L, in_ch, out_ch):
    from upliftingtabletennis_b200.detector import HRNetEngine
    e = HRNetEngine(in_ch, out_ch, 0, out_ch)
    ref = [(s.name, s.bn, s.cin, s.cout, s.k, s.stride) for s in ohr.conv_specs(in_ch, out_ch)]
    assert e.specs == ref
    assert len(e.specs) == 72
    # workspace planning is host-side arithmetic: a 1280x704 stack in bf16 must fit comfortably in HBM
    ws = L.lib.ttk_hrnet_workspace_bytes(e.h, 1, 704, 1280, L.BF16)
    assert 100e6 < ws < 4e9, ws
    assert L.lib.ttk_hrnet_workspace_bytes(e.h, 1, 704, 1280, L.F32) == pytest.approx(2 * ws, rel=0.01)


def test_uplift_params_match_reference_layout(L):
    from upliftingtabletennis_b200.uplift import UpliftEngine
    e = UpliftEngine(128, 4, 16, True)
    ref = oup.state_dict_layout('large')
    assert [n for n, _ in e.params] == [n for n, _ in ref]
    for (n, numel), (_, shape) in zip(e.params, ref):
        k = 1
        for s in shape:
            k *= s
        assert numel == k, n


def test_module_shells_load_reference_state_dicts():
    from upliftingtabletennis_b200.detector import MyHRNet, WASBNet
    from upliftingtabletennis_b200.uplift import get_model
    assert not WASBNet().load_state_dict(ohr.random_state_dict(9, 3, 1), strict=True).missing_keys
    assert not MyHRNet().load_state_dict(ohr.random_state_dict(3, 13, 1), strict=True).missing_keys
    for name in ('connectstage', 'multistage'):
        m = get_model(name, 'large', 'dynamic', 'new')
        assert not m.load_state_dict(oup.random_state_dict(2), strict=True).missing_keys
    with pytest.raises(NotImplementedError):
        get_model('singlestage', 'large', 'dynamic', 'new')


def test_argument_validation_without_gpu(L):
    lib = L.lib
    h = C.c_void_p()
    assert lib.ttk_hrnet_create(99, 3, 1, 1, C.byref(h)) == -1
    assert b'in_ch' in lib.ttk_last_error()
    assert lib.ttk_uplift_create(64, 4, 12, 1, C.byref(h)) == -4          # only the 'large' model has kernels
    assert lib.ttk_heatmap_decode(None, 4, 0, 8, 0, 1920, 1080, None, None, None, None, 0, None) == -1
    assert lib.ttk_heatmap_decode(None, 4, 8, 8, 7, 1920, 1080, None, None, None, None, 0, None) == -1
    assert lib.ttk_heatmap_decode(None, 0, 8, 8, 0, 1920, 1080, None, None, None, None, 0, None) == 0   # empty batch is fine
    assert lib.ttk_preprocess_stacks(None, 3, 8, 8, 2, 1, 1, 8, 8, None, None, 0, 0, None) == -1
    assert lib.ttk_decode_workspace_bytes(32, 704, 1280) > 0


def test_no_cpu_fallback(L):
    import torch
    from upliftingtabletennis_b200.detector import WASBNet
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        WASBNet()(torch.zeros(1, 9, 8, 8))
    with pytest.raises(L.TtkError):
        L.require_device()
