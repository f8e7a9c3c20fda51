"""ViTPose-small detector (SURVEY.md section 8 row a4'): CUDA path against the reference's golden outputs and the oracle."""
import numpy as np
import pytest
import torch

from oracle import vitpose as ov

pytestmark = pytest.mark.gpu
TOL = dict(rtol=0.0, atol=1e-4)      # |d| <= 1e-4 max|h| + 1e-5, applied through `close`


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    from upliftingtabletennis_b200 import _lib
    _lib.require_device()
    return torch.device('cuda:0')


def close(y, ref, rel=1e-4):
    bound = rel * np.abs(ref).max() + 1e-5
    err = np.abs(y - ref).max()
    assert err <= bound, (err, bound)


def test_vitpose_fp32_golden(dev, golden):
    from upliftingtabletennis_b200.vitpose import TableVitPose, VitPose
    g = golden('vitpose')
    m = VitPose(in_frames=3, model_size='small', resolution=(96, 64), dtype='fp32').to(dev).eval()
    m.load_state_dict(ov.random_state_dict(int(g['ball_seed']), 9, 24, 1), strict=True)
    y, none = m(torch.from_numpy(g['ball_x']).to(dev))
    assert none is None and tuple(y.shape) == g['ball_y'].shape
    close(y.cpu().numpy(), g['ball_y'])
    t = TableVitPose(model_size='small', resolution=(96, 64), dtype='fp32').to(dev).eval()
    t.load_state_dict(ov.random_state_dict(int(g['table_seed']), 3, 24, 13), strict=True)
    yt = t(torch.from_numpy(g['table_x']).to(dev))
    close(yt.cpu().numpy(), g['table_y'])


@pytest.mark.parametrize('res,batch', [((160, 96), 3), ((1152, 640), 1)])
def test_vitpose_fp32_vs_oracle(dev, res, batch):
    """Other token grids (odd batch, tokens not a multiple of the tile sizes) and the full 1152 x 640 configuration."""
    from upliftingtabletennis_b200.vitpose import VitPose
    hp, wp = ov.tokens_hw(res[1], res[0])
    sd = ov.random_state_dict(7, 9, hp * wp, 1)
    m = VitPose(in_frames=3, resolution=res, dtype='fp32').to(dev).eval()
    m.load_state_dict(sd, strict=True)
    x = np.random.default_rng(1).standard_normal((batch, 9, res[1], res[0])).astype(np.float32)
    y, _ = m(torch.from_numpy(x).to(dev))
    ref = ov.vitpose_forward(sd, x).numpy()
    assert tuple(y.shape) == (batch, 1, 4 * hp, 4 * wp)
    close(y.cpu().numpy(), ref)
    # decoded peak identical where the oracle's top-2 margin exceeds twice the tolerance
    for b in range(batch):
        flat = ref[b, 0].ravel()
        top2 = np.partition(flat, -2)[-2:]
        if top2[1] - top2[0] > 2 * (1e-4 * np.abs(ref).max() + 1e-5):
            assert int(y[b, 0].flatten().argmax()) == int(flat.argmax())


def _split(t):
    """tf32 hi / lo pair of a float32 tensor as the library's kernels produce it (cvt.rna: round to nearest, ties away)."""
    def rna(v):
        b = v.contiguous().view(torch.int32)
        return ((b + 0x1000) & ~0x1FFF).view(torch.float32)
    hi = rna(t)
    return torch.stack([hi, rna(t - hi)]).contiguous()


@pytest.mark.parametrize('M,N,K,act,res,split', [
    (2880, 384, 2304, 0, True, False),       # patch embedding (+ position table)
    (5760, 1152, 384, 0, False, True),       # qkv -> split pair for the attention
    (5760, 1536, 384, 1, False, True),       # fc1 + GELU -> split pair for fc2
    (5760, 384, 1536, 0, True, False),       # fc2 + residual
    (300, 384, 384, 0, True, False),         # partial M tile
    (24, 1152, 384, 0, False, True),         # fewer rows than one tile
    (130, 128, 32, 2, False, False),         # single K block
])
def test_gemm_x3(dev, M, N, K, act, res, split):
    """3xTF32 GEMM (gemm_umma.cu X3 mode) against float64: float32-class error -- the split leaves ~2^-21 per product, the rest is the
    tensor core's float32 accumulation over K (measured 6.5e-6 of max|C| at K = 2304) -- two orders below one TF32 product's 5e-4."""
    from upliftingtabletennis_b200 import _lib
    from upliftingtabletennis_b200._lib import lib, check, ptr, stream_ptr
    g = torch.Generator(device='cpu').manual_seed(M + N)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    R = torch.randn(M, N, generator=g).to(dev) if res else None
    ref = A.double() @ W.double().t() + bias.double()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    if act == 2:
        ref = torch.relu(ref)
    if res:
        ref = ref + R.double()
    A2, W2 = _split(A), _split(W)
    C = torch.full((2 if split else 1, M, N), float('nan'), dtype=torch.float32, device=dev)
    check(lib.ttk_vit_debug_gemm(ptr(A2), ptr(W2), ptr(bias), ptr(R) if res else None, ptr(C), M, N, K, act, 2 if split else 0, 0, 0, 0, 0,
                                 _lib.TF32X3, stream_ptr()))
    torch.cuda.synchronize()
    out = C.double().sum(0)
    err = (out - ref).abs().max().item()
    assert err <= 1.5e-5 * ref.abs().max().item(), (M, N, K, err, ref.abs().max().item())
    if split:                                 # the hi plane is a TF32 number, the lo plane what is left of the float32 value
        assert torch.equal(C[0], _split(C[0])[0])
        assert (C[1].abs() <= C[0].abs() * 2.0 ** -10 + 1e-30).all()


def test_gemm_x3_deconv_implicit(dev):
    """Transposed convolution (4x4, stride 2) as four implicit GEMMs: 5-D TMA gather of the 2x2 taps from the split NHWC feature map."""
    from upliftingtabletennis_b200.vitpose import VitPose
    # exercised through the whole detector below (the debug hook has no implicit mode); here only a smoke check that a
    # resolution whose token grid is not a multiple of the 16 x 8 pixel patch runs through the head
    hp, wp = ov.tokens_hw(64, 96)
    m = VitPose(in_frames=3, resolution=(96, 64), dtype='tf32x3').to(dev).eval()
    m.load_state_dict(ov.random_state_dict(3, 9, hp * wp, 1), strict=True)
    y, _ = m(torch.randn(1, 9, 64, 96, device=dev))
    assert torch.isfinite(y).all()


@pytest.mark.parametrize('images,tokens', [(1, 2880), (2, 24), (3, 60), (2, 128), (1, 300)])
def test_attention_x3(dev, images, tokens):
    from upliftingtabletennis_b200 import _lib
    from upliftingtabletennis_b200._lib import lib, check, ptr, stream_ptr
    g = torch.Generator(device='cpu').manual_seed(tokens)
    qkv = (torch.randn(images * tokens, 1152, generator=g) * 1.5).to(dev)
    out = torch.full((2, images * tokens, 384), float('nan'), dtype=torch.float32, device=dev)
    scratch = torch.empty((2 * images * 12 * 48 * (tokens + 8) * 4,), dtype=torch.uint8, device=dev)
    check(lib.ttk_vit_debug_attention(ptr(_split(qkv)), ptr(out), ptr(scratch), scratch.numel(), images, tokens, _lib.TF32X3, stream_ptr()))
    torch.cuda.synchronize()
    q, k, v = qkv.double().view(images, tokens, 3, 12, 32).permute(2, 0, 3, 1, 4)
    ref = torch.softmax((q * 32 ** -0.5) @ k.transpose(-2, -1), dim=-1) @ v
    ref = ref.transpose(1, 2).reshape(images * tokens, 384)
    err = (out.double().sum(0) - ref).abs().max().item()
    assert err <= 4e-6 * ref.abs().max().item() + 1e-6, (err, ref.abs().max().item())


@pytest.mark.parametrize('res,batch', [((96, 64), 2), ((160, 96), 3), ((1152, 640), 1)])
def test_vitpose_x3_vs_oracle(dev, res, batch):
    """The whole detector in the tf32x3 class meets the float32 parity bound (|d| <= 1e-4 max|h| + 1e-5) like the SIMT path."""
    from upliftingtabletennis_b200.vitpose import TableVitPose, VitPose
    hp, wp = ov.tokens_hw(res[1], res[0])
    sd = ov.random_state_dict(7, 9, hp * wp, 1)
    m = VitPose(in_frames=3, resolution=res, dtype='tf32x3').to(dev).eval()
    m.load_state_dict(sd, strict=True)
    x = np.random.default_rng(1).standard_normal((batch, 9, res[1], res[0])).astype(np.float32)
    y, _ = m(torch.from_numpy(x).to(dev))
    ref = ov.vitpose_forward(sd, x).numpy()
    close(y.cpu().numpy(), ref)
    for b in range(batch):
        flat = ref[b, 0].ravel()
        top2 = np.partition(flat, -2)[-2:]
        if top2[1] - top2[0] > 2 * (1e-4 * np.abs(ref).max() + 1e-5):
            assert int(y[b, 0].flatten().argmax()) == int(flat.argmax())
    if res == (96, 64):                       # 13-channel table variant
        sdt = ov.random_state_dict(11, 3, hp * wp, 13)
        t = TableVitPose(resolution=res, dtype='tf32x3').to(dev).eval()
        t.load_state_dict(sdt, strict=True)
        xt = np.random.default_rng(2).standard_normal((batch, 3, res[1], res[0])).astype(np.float32)
        close(t(torch.from_numpy(xt).to(dev)).cpu().numpy(), ov.vitpose_forward(sdt, xt).numpy())


# ------------------------------------------------------------------------------------------------
# bf16 tensor-core path: the two tcgen05 kernels against torch on the same bf16 operands, then the whole detector
# ------------------------------------------------------------------------------------------------
def _gemm(dev, M, N, K, act, res, c_bf16, up=None, seed=0):
    from upliftingtabletennis_b200 import _lib
    from upliftingtabletennis_b200._lib import lib, check, ptr, stream_ptr
    g = torch.Generator(device='cpu').manual_seed(seed)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    R = torch.randn(M, N, generator=g).to(dev) if res else None
    ref = A.float() @ W.float().t() + bias
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    if act == 2:
        ref = torch.relu(ref)
    if res:
        ref = ref + R
    rows = M
    up_h = up_w = py = px = 0
    if up:
        up_h, up_w, py, px = up
        rows = 4 * M
    C = torch.full((rows, N), float('nan'), dtype=torch.bfloat16 if c_bf16 else torch.float32, device=dev)
    check(lib.ttk_vit_debug_gemm(ptr(A), ptr(W), ptr(bias), ptr(R) if res else None, ptr(C), M, N, K, act, int(c_bf16), up_h, up_w, py, px,
                                 _lib.BF16, stream_ptr()))
    torch.cuda.synchronize()
    if up:
        m = torch.arange(M, device=dev)
        x, y, img = m % up_w, (m // up_w) % up_h, m // (up_w * up_h)
        idx = (img * 2 * up_h + 2 * y + py) * 2 * up_w + 2 * x + px
        out = C[idx]
        mask = torch.ones(rows, dtype=torch.bool, device=dev)
        mask[idx] = False
        assert torch.isnan(C[mask].float()).all()           # rows of the other parities are untouched
    else:
        out = C
    err = (out.float() - ref).abs().max().item()
    tol = (2e-2 if c_bf16 else 2e-3) * ref.abs().max().item()
    assert err <= tol, (M, N, K, err, tol)


@pytest.mark.parametrize('M,N,K,act,res,c_bf16', [
    (2880, 384, 2304, 0, True, False),       # patch embedding of one image (+ position table as residual)
    (5760, 1152, 384, 0, False, True),       # qkv
    (5760, 1536, 384, 1, False, True),       # fc1 + GELU
    (5760, 384, 1536, 0, True, False),       # fc2 + residual, float32 stream
    (5760, 384, 384, 0, True, False),        # proj + residual (weights resident in shared memory, like qkv / fc1 at this M)
    (2100, 1152, 384, 0, False, True),       # resident variant with a partial last M tile and idle CTAs
    (300, 384, 384, 0, True, False),         # partial M tile
    (24, 1152, 384, 0, False, True),         # fewer rows than one tile
    (130, 128, 64, 2, False, True),          # BN = 128 path, single K block
])
def test_gemm_umma(dev, M, N, K, act, res, c_bf16):
    _gemm(dev, M, N, K, act, res, c_bf16)


def test_gemm_umma_deconv_scatter(dev):
    _gemm(dev, 2 * 40 * 72, 256, 1536, 2, False, True, up=(40, 72, 1, 0))
    _gemm(dev, 5 * 7 * 3, 256, 1024, 2, False, True, up=(5, 7, 0, 1))


@pytest.mark.parametrize('images,tokens', [(1, 2880), (2, 24), (3, 60), (2, 128), (1, 300)])
def test_attention_umma(dev, images, tokens):
    from upliftingtabletennis_b200 import _lib
    from upliftingtabletennis_b200._lib import lib, check, ptr, stream_ptr
    g = torch.Generator(device='cpu').manual_seed(tokens)
    qkv = (torch.randn(images * tokens, 1152, generator=g) * 1.5).to(torch.bfloat16).to(dev)
    out = torch.full((images * tokens, 384), float('nan'), dtype=torch.bfloat16, device=dev)
    scratch = torch.empty((images * 12 * 48 * (tokens + 8) * 2,), dtype=torch.uint8, device=dev)       # V^T + ones row per head
    check(lib.ttk_vit_debug_attention(ptr(qkv), ptr(out), ptr(scratch), scratch.numel(), images, tokens, _lib.BF16, stream_ptr()))
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(images, tokens, 3, 12, 32).permute(2, 0, 3, 1, 4)
    ref = torch.softmax((q * 32 ** -0.5) @ k.transpose(-2, -1), dim=-1) @ v
    ref = ref.transpose(1, 2).reshape(images * tokens, 384)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, (err, ref.abs().max().item())
    # and the float32 kernel against the same reference
    out32 = torch.empty((images * tokens, 384), dtype=torch.float32, device=dev)
    check(lib.ttk_vit_debug_attention(ptr(qkv.float().contiguous()), ptr(out32), None, 0, images, tokens, _lib.F32, stream_ptr()))
    assert (out32 - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-5


def test_vitpose_bf16_bound(dev, golden):
    """bf16 tensor-core path, reported separately: relative L2 error of the heatmap against the float32 oracle."""
    from upliftingtabletennis_b200.vitpose import VitPose
    for res, batch in (((96, 64), 2), ((1152, 640), 1)):
        hp, wp = ov.tokens_hw(res[1], res[0])
        sd = ov.random_state_dict(7, 9, hp * wp, 1)
        m = VitPose(in_frames=3, resolution=res, dtype='fp32').to(dev).eval()
        m.load_state_dict(sd, strict=True)
        x = np.random.default_rng(1).standard_normal((batch, 9, res[1], res[0])).astype(np.float32)
        y32, _ = m(torch.from_numpy(x).to(dev))
        m.compute_dtype = torch.bfloat16
        y16, _ = m(torch.from_numpy(x).to(dev))
        assert torch.isfinite(y16).all()
        rel = ((y16 - y32).norm() / y32.norm()).item()
        assert rel < 3e-2, (res, rel)
