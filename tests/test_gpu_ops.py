"""GPU parity of the tensor-core / helper operators (through the C ABI) against plain PyTorch fp32
references on the same 16-bit-rounded inputs (both operand formats: fp16 and bf16)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


H16 = [torch.float16, torch.bfloat16]
# output rounding of the 16-bit result: 2^-11 (fp16) / 2^-8 (bf16) relative, plus accumulation-order noise
TOL = {torch.float16: 2e-3, torch.bfloat16: 1e-2}


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def _relerr(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6))


@pytest.mark.parametrize("M,N,K,act", [
    (128, 256, 64, 0), (300, 384, 384, 0), (1370, 1152, 384, 0), (2740, 3072, 1024, 1), (129, 48, 384, 0),
    (1369, 96, 768, 2), (640, 1024, 640, 0), (257, 32, 128, 0), (2 * 1370, 768, 3072, 0)])
@pytest.mark.parametrize("dt", H16)
def test_linear_h16(M, N, K, act, dt):
    from dav2_b200 import ops
    a, w = _rand((M, K), 1, 1.0, dt), _rand((N, K), 2, K ** -0.5, dt)
    bias = _rand((N,), 3, 0.1, torch.float32)
    out = ops.linear_h16(a, w, bias, act)
    ref = a.float() @ w.float().t() + bias
    ref = F.gelu(ref) if act == 1 else (F.relu(ref) if act == 2 else ref)
    assert out.shape == (M, N) and out.dtype == dt
    assert _relerr(out, ref) < TOL[dt]


@pytest.mark.parametrize("M,N,K", [(200, 384, 384), (1370, 1024, 4096), (2741, 768, 768)])
@pytest.mark.parametrize("dt", H16)
def test_linear_resid(M, N, K, dt):
    from dav2_b200 import ops
    a, w = _rand((M, K), 4, 1.0, dt), _rand((N, K), 5, K ** -0.5, dt)
    bias, gamma = _rand((N,), 6, 0.1, torch.float32), _rand((N,), 7, 1.0, torch.float32)
    x = _rand((M, N), 8, 1.0, torch.float32)
    ref = x + gamma * (a.float() @ w.float().t() + bias)
    ops.linear_resid_(x, a, w, bias, gamma)
    assert _relerr(x, ref) < 2e-5  # fp32 accumulate + fp32 residual


@pytest.mark.parametrize("B,H,W,Cin,Cout,act", [
    (1, 8, 16, 64, 64, 0), (2, 19, 19, 256, 256, 2), (1, 37, 37, 48, 64, 0), (2, 74, 74, 128, 128, 0),
    (1, 50, 45, 64, 32, 0), (1, 37, 41, 1024, 256, 0), (2, 148, 148, 64, 64, 2)])
@pytest.mark.parametrize("dt", H16)
def test_conv3x3(B, H, W, Cin, Cout, act, dt):
    from dav2_b200 import ops
    x = _rand((B, H, W, Cin), 9, 1.0, dt)
    w = _rand((Cout, Cin, 3, 3), 10, (9 * Cin) ** -0.5, dt)
    bias = _rand((Cout,), 11, 0.1, torch.float32)
    add1, add2 = _rand((B, H, W, Cout), 12, 1.0, dt), _rand((B, H, W, Cout), 13, 1.0, dt)
    out, out_relu = ops.conv3x3_h16(x, ops.pack_conv3x3_weight(w), bias, add1, add2, act, want_relu=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
    ref = F.relu(ref) if act == 2 else ref
    ref = ref.permute(0, 2, 3, 1) + add1.float() + add2.float()
    assert _relerr(out, ref) < TOL[dt]
    assert _relerr(out_relu, F.relu(ref)) < TOL[dt]


@pytest.mark.parametrize("B,N,heads", [(1, 128, 1), (2, 50, 6), (1, 1370, 6), (2, 300, 16), (1, 5477, 2)])
@pytest.mark.parametrize("dt", H16)
def test_attention(B, N, heads, dt):
    from dav2_b200 import ops
    D = heads * 64
    qkv = _rand((B * N, 3 * D), 14, 1.0, dt)
    qkv[:, :D] *= 0.125  # q arrives pre-scaled by d^-1/2 (folded into the qkv weights)
    qkv = qkv.contiguous()
    out = ops.attention_h16(qkv, B, N, D)
    q, k, v = (t.reshape(B, N, heads, 64).transpose(1, 2).float() for t in qkv.split(D, dim=1))
    ref = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).transpose(1, 2).reshape(B * N, D)
    assert _relerr(out, ref) < 2 * TOL[dt]  # P is rounded to 16 bits before the PV product


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("case", ["late_peak", "ramp", "early_peak", "huge", "tail_peak"])
def test_attention_reference_max_slow_path(case, dt):
    """The attention kernel keeps a REFERENCE max per row and rescales the TMEM accumulator lazily (only when a row's max
    grew by more than 2^8); its -DATTN_SPECULATIVE build skips the per-tile max altogether and recentres when a tile's row
    sum exceeds 2^13.  These score patterns force both slow paths: a key far above the earlier maxima late in the sequence,
    scores ramping up tile after tile, the mirror image (later tiles underflow), scores in the hundreds, and a peak inside
    the masked tail tile."""
    from dav2_b200 import ops
    B, N, heads = 2, 300, 2
    D = heads * 64
    g = torch.Generator().manual_seed(31)
    q = torch.randn(B, N, heads, 64, generator=g) * 0.125
    k = torch.randn(B, N, heads, 64, generator=g)
    v = torch.randn(B, N, heads, 64, generator=g)
    u = torch.randn(64, generator=g)
    u = u / u.norm()
    q = q + 1.0 * u  # every query has a component along u: adding c*u to a key adds ~c to its score
    pos = torch.arange(N).view(1, N, 1, 1).float()
    if case == "late_peak":
        k[:, 200] += 30.0 * u          # score +30 (43 octaves) at key 200: tile 3 of 5
    elif case == "ramp":
        k = k + (pos / N) * 60.0 * u   # +12 per tile
    elif case == "early_peak":
        k[:, 3] += 40.0 * u            # the first tile holds the max: every later tile underflows
    elif case == "huge":
        k = k + (pos % 7 - 3) * 40.0 * u
    else:
        k[:, N - 1] += 50.0 * u        # last valid key of the (masked) tail tile
    qkv = torch.cat([q.reshape(B * N, D), k.reshape(B * N, D), v.reshape(B * N, D)], dim=1).to(dt).cuda().contiguous()
    out = ops.attention_h16(qkv, B, N, D)
    qf, kf, vf = (t.reshape(B, N, heads, 64).transpose(1, 2).float() for t in qkv.split(D, dim=1))
    ref = (torch.softmax(qf @ kf.transpose(-1, -2), dim=-1) @ vf).transpose(1, 2).reshape(B * N, D)
    assert torch.isfinite(out.float()).all()
    assert _relerr(out, ref) < 2 * TOL[dt]


@pytest.mark.parametrize("rows,D", [(5, 384), (1370, 768), (2741, 1024)])
@pytest.mark.parametrize("dt", H16)
def test_layernorm(rows, D, dt):
    from dav2_b200 import ops
    x = _rand((rows, D), 15, 3.0, torch.float32) + 0.5
    w, b = _rand((D,), 16, 1.0, torch.float32), _rand((D,), 17, 1.0, torch.float32)
    out = ops.layernorm(x, w, b, 1e-6, dt)
    assert _relerr(out, F.layer_norm(x, (D,), w, b, 1e-6)) < TOL[dt]


@pytest.mark.parametrize("B,Hi,Wi,Ho,Wo,C", [(1, 19, 19, 37, 37, 64), (2, 37, 37, 74, 74, 256), (1, 296, 296, 518, 518, 32),
                                              (1, 37, 78, 74, 156, 128)])
@pytest.mark.parametrize("dt", H16)
def test_bilinear(B, Hi, Wi, Ho, Wo, C, dt):
    from dav2_b200 import ops
    x = _rand((B, Hi, Wi, C), 18, 1.0, dt)
    out = ops.bilinear_nhwc_h16(x, Ho, Wo)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), (Ho, Wo), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    assert _relerr(out, ref) < TOL[dt]
    d = _rand((B, Hi, Wi), 19, 1.0, torch.float32)
    r = F.interpolate(d[:, None], (Ho, Wo), mode="bilinear", align_corners=True)[:, 0]
    assert _relerr(ops.resize_depth(d, Ho, Wo), r) < 1e-5
