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


def test_precision_classes_and_defaults(L):
    """The arithmetic classes (precision.py) and which one every model shell runs by default: the measured tensor-core paths."""
    import torch
    from upliftingtabletennis_b200 import precision as P
    from upliftingtabletennis_b200.detector import MyHRNet, WASBNet
    from upliftingtabletennis_b200.uplift import get_model
    from upliftingtabletennis_b200.vitpose import VitPose
    assert (L.F32, L.BF16, L.TF32, L.TF32X3) == (0, 1, 2, 3)
    header = open(os.path.join(ROOT, 'include', 'ttk.h')).read()
    assert 'TTK_F32 = 0, TTK_BF16 = 1, TTK_TF32 = 2, TTK_TF32X3 = 3' in header
    assert P.canonical('TF32') == 'tf32' and P.canonical(torch.float32) == 'fp32' and P.canonical(torch.bfloat16) == 'bf16' and P.canonical('3xtf32') == 'tf32x3'
    assert P.storage_dtype('tf32') == torch.float32 and P.storage_dtype('tf32x3') == torch.float32 and P.storage_dtype('bf16') == torch.bfloat16
    with pytest.raises(ValueError):
        P.canonical('fp8')
    assert WASBNet().compute_dtype == 'tf32' and MyHRNet().compute_dtype == 'tf32'
    assert WASBNet(dtype=torch.bfloat16).compute_dtype == 'bf16' and WASBNet(dtype='fp32').storage_dtype == torch.float32
    up = get_model('connectstage', 'large', 'dynamic', 'new')
    assert up.compute_dtype == 'tf32x3'
    up.compute_dtype = torch.float32
    assert up.compute_dtype == 'fp32'
    with pytest.raises(NotImplementedError):
        up.compute_dtype = 'tf32'                      # the transformer has no single-product TF32 path
    assert VitPose(in_frames=3, resolution=(96, 64)).compute_dtype == 'tf32x3'
    with pytest.raises(NotImplementedError):
        VitPose(in_frames=3, resolution=(96, 64), dtype='tf32')
    # workspace of the tf32x3 path: a layer's activations live in HBM (x, q | k | v, o per table-token row)
    from upliftingtabletennis_b200.uplift import UpliftEngine
    e = UpliftEngine(128, 4, 16, True)
    base = L.lib.ttk_uplift_workspace_bytes(e.h, 8, 50, L.F32)
    assert L.lib.ttk_uplift_workspace_bytes(e.h, 8, 50, L.TF32X3) - base >= 8 * 50 * 14 * (128 + 384 + 128) * 4


def test_reference_seams_importable_and_fail_early():
    """The reference's module paths (inference/...) and the hub entry points; unsupported model names fail before any download."""
    import inspect
    import hubconf
    import inference.inference_balldetection as ib
    import inference.inference_tabledetection as it
    import inference.inference_uplifting as iu
    from inference import utils
    from upliftingtabletennis_b200 import interface
    assert ib.load_model is interface.load_ball_model and it.load_model is interface.load_table_model and iu.load_model is interface.load_uplifting_model
    assert list(inspect.signature(utils.process_trajectory_ball).parameters) == ['ball_model', 'images', 'move_weights']
    assert list(inspect.signature(utils.process_trajectory_table).parameters) == ['table_model', 'images', 'move_weights']
    assert list(inspect.signature(utils.process_trajectory_uplifting).parameters) == [
        'uplifting_model', 'predictions_ball', 'predictions_table', 'times', 'mask', 'transform_mode', 'move_weights']
    assert list(inspect.signature(utils.extract_position_table).parameters) == ['heatmaps', 'image_width', 'image_height', 'threshold']
    assert hubconf.dependencies == ['torch', 'numpy']
    for fn in (hubconf.ball_detection, hubconf.table_detection):           # the reference's default name is kept and fails with a clear error
        assert inspect.signature(fn).parameters['model_name'].default == 'segformerpp_b2'
        with pytest.raises(NotImplementedError):
            fn()
    sig = inspect.signature(interface.TableTennisPipeline.__init__).parameters
    assert (sig['ball_model'].default, sig['ball_model_aux'].default, sig['table_model'].default, sig['table_model_aux'].default) == (
        'vitpose', 'wasb', 'vitpose', 'hrnet')


def test_hrnet_plan_drops_unread_outputs(L):
    """Dead-code elimination of the plan is host logic: the reference graph has 72 convolutions, 13 of them (fuse outputs 1-3 of stage 4)
    feed nothing; a forward pass therefore launches 72 - 13 - 1 (projection shortcut folded into conv3) conv kernels + 3 sums + 1 final."""
    from upliftingtabletennis_b200.detector import HRNetEngine
    e = HRNetEngine(9, 3, 1, 1)
    names = [s[0] for s in e.specs]
    dead = [n for n in names if n.startswith('model.stage4.0.fuse_layers.') and not n.startswith('model.stage4.0.fuse_layers.0.')]
    assert len(dead) == 13
