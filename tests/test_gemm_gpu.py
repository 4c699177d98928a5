"""GPU parity of the tcgen05 GEMM / implicit-GEMM conv kernel (through the C ABI) against plain
torch fp32 ops on identically pre-rounded 16-bit inputs. Tolerances: fp32 accumulation of exactly
representable products -> 1e-4 relative to the output scale for fp32 outputs, one 16-bit ulp otherwise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DT = [torch.float16, torch.bfloat16]


def _ops():
    from l4p_b200 import ops
    return ops


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def _close(got, ref, tol):
    scale = ref.abs().max().item() + 1e-12
    err = (got.float() - ref.float()).abs().max().item()
    assert err <= tol * scale, f"max abs err {err:.3e} > {tol:.1e} * {scale:.3e}"


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (2048, 1408, 1408), (384, 176, 1176),
                                   (2048, 4224, 1408), (130, 48, 72), (2048, 6144, 1408)])
def test_linear_plain(dtype, M, N, K):
    ops = _ops()
    a = _rand((M, K), dtype, 1)
    w = _rand((N, K), dtype, 2, K ** -0.5)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.linear(a, w, out_f32=out)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    _close(out, ref, 2e-5)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M,N,K,pair,bn", [(4096, 384, 704, 1, 0),      # 2-CTA kernel, 11 k-blocks: odd tail stage (one valid slab)
                                           (4096, 384, 448, 1, 128),    # 7 k-blocks, forced 128-wide tiles
                                           (512, 64, 320, -1, 64),      # 1-CTA kernel, 5 k-blocks, three 128-wide stages
                                           (640, 96, 1408, -1, 96),     # 1-CTA kernel, 22 k-blocks (even)
                                           (2048, 1408, 6144, 0, 0),    # fc2 as shipped (9 x 160, one round of 72 pair tiles)
                                           (300, 208, 256, 1, 0),       # ragged M, 4 k-blocks = two full stages
                                           (2048, 1408, 264, 0, 0)])    # K % 64 != 0: 64-wide stages (plain 2-D boxes, zero-filled tail)
def test_linear_k128_stages(dtype, M, N, K, pair, bn):
    """Matrix mode with K % 64 == 0 runs ring stages of two k-blocks per operand (3-D TMA boxes, slab-major in shared memory);
    odd k-block counts end on a half-filled stage whose second slab is skipped. Columns are scaled per k-block so that a wrong
    slab pairing (A slab i against B slab j) cannot cancel out."""
    ops = _ops()
    a = _rand((M, K), dtype, 31)
    w = _rand((N, K), dtype, 32, K ** -0.5)
    scale = (1.0 + 0.25 * (torch.arange(K, device="cuda") // 64)).to(dtype)
    a = a * scale
    b = _rand((N,), torch.float32, 33)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.linear(a, w, bias=b, out_f32=out, cta_pair=pair, block_n=bn)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + b
    _close(out, ref, 2e-5)


@pytest.mark.parametrize("dtype", DT)
def test_linear_bias_gelu_16(dtype):
    ops = _ops()
    from l4p_b200 import lib
    M, N, K = 512, 6144, 1408
    a = _rand((M, K), dtype, 3)
    w = _rand((N, K), dtype, 4, K ** -0.5)
    b = _rand((N,), torch.float32, 5)
    out = torch.empty(M, N, device="cuda", dtype=dtype)
    ops.linear(a, w, bias=b, act=lib.ACT_GELU, out_16=out)
    torch.cuda.synchronize()
    ref = F.gelu(a.float() @ w.float().t() + b)
    _close(out, ref, 2 ** -8 if dtype == torch.bfloat16 else 2 ** -10)


@pytest.mark.parametrize("dtype", DT)
def test_linear_residual_inplace(dtype):
    ops = _ops()
    M, N, K = 2048, 1408, 6144
    a = _rand((M, K), dtype, 6)
    w = _rand((N, K), dtype, 7, K ** -0.5)
    b = _rand((N,), torch.float32, 8)
    x = _rand((M, N), torch.float32, 9)
    ref = x + a.float() @ w.float().t() + b
    o16 = torch.empty(M, N, device="cuda", dtype=dtype)
    ops.linear(a, w, bias=b, res_f32=x, out_f32=x, out_16=o16)
    torch.cuda.synchronize()
    _close(x, ref, 2e-5)
    _close(o16, ref, 2 ** -8 if dtype == torch.bfloat16 else 2 ** -10)


@pytest.mark.parametrize("dtype", DT)
def test_linear_qkv_scatter(dtype):
    ops = _ops()
    B, H, d, dp, Ntok, Cdim = 2, 16, 88, 96, 256, 1408
    a = _rand((B * Ntok, Cdim), dtype, 10)
    w = _rand((3 * H * d, Cdim), dtype, 11, Cdim ** -0.5)
    b = _rand((3 * H * d,), torch.float32, 12)
    q = torch.zeros(B, H, Ntok, dp, device="cuda", dtype=dtype)
    k = torch.zeros_like(q)
    vt = torch.zeros(B, H, dp, Ntok, device="cuda", dtype=dtype)
    ops.linear_qkv(a, w, b, q, k, vt, H, d, Ntok)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t() + b).reshape(B, Ntok, 3, H, d).permute(2, 0, 3, 1, 4)
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -10
    _close(q[..., :d], ref[0], tol)
    _close(k[..., :d], ref[1], tol)
    _close(vt[:, :, :d].transpose(-1, -2), ref[2], tol)
    assert q[..., d:].abs().max().item() == 0 and vt[:, :, d:].abs().max().item() == 0


def _conv_ref(x_cl, w_k, ksize, bias=None):
    """x_cl [B,T,H,W,C]; w_k [Cout, taps*Cin] (kt,kh,kw,cin) -> [B,T,H,W,Cout] fp32."""
    B, T, H, W, Cin = x_cl.shape
    kT, kH, kW = ksize
    Cout = w_k.shape[0]
    w = w_k.float().reshape(Cout, kT, kH, kW, Cin).permute(0, 4, 1, 2, 3).contiguous()
    y = F.conv3d(x_cl.float().permute(0, 4, 1, 2, 3), w, bias, padding=(kT // 2, kH // 2, kW // 2))
    return y.permute(0, 2, 3, 4, 1).contiguous()


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(1, 4, 8, 8, 64, 64), (2, 8, 16, 16, 128, 256), (1, 2, 32, 32, 256, 128),
                                   (1, 3, 20, 24, 64, 32)])
def test_conv3d_3x3x3(dtype, shape):
    ops = _ops()
    B, T, H, W, Cin, Cout = shape
    x = _rand((B, T, H, W, Cin), dtype, 13)
    w = _rand((Cout, 27 * Cin), dtype, 14, (27 * Cin) ** -0.5)
    b = _rand((Cout,), torch.float32, 15)
    out = torch.empty(B, T, H, W, Cout, device="cuda", dtype=torch.float32)
    ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, out_f32=out)
    torch.cuda.synchronize()
    _close(out, _conv_ref(x, w, (3, 3, 3), b), 3e-5)


@pytest.mark.parametrize("dtype", DT)
def test_conv3d_rcu_epilogue(dtype):
    """conv + bias + two 16-bit residuals, dual output (raw and ReLU'd) as used by the RefineNet blocks."""
    ops = _ops()
    B, T, H, W, Cc = 1, 4, 16, 16, 256
    x = _rand((B, T, H, W, Cc), dtype, 16)
    w = _rand((Cc, 27 * Cc), dtype, 17, (27 * Cc) ** -0.5)
    b = _rand((Cc,), torch.float32, 18)
    r1 = _rand((B, T, H, W, Cc), dtype, 19)
    r2 = _rand((B, T, H, W, Cc), dtype, 20)
    o = torch.empty_like(x)
    orl = torch.empty_like(x)
    ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, res_16=r1, res2_16=r2, out_16=o, out_16_relu=orl)
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, (3, 3, 3), b) + r1.float() + r2.float()
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -10
    _close(o, ref, tol)
    _close(orl, ref.clamp_min(0), tol)


@pytest.mark.parametrize("dtype", DT)
def test_conv3d_head1x1(dtype):
    ops = _ops()
    B, T, H, W, Cc, C2 = 1, 2, 28, 32, 128, 2
    x = _rand((B, T, H, W, Cc), dtype, 21)
    w = _rand((Cc, 27 * Cc), dtype, 22, (27 * Cc) ** -0.5)
    b = _rand((Cc,), torch.float32, 23)
    w2 = _rand((C2, Cc), torch.float32, 24, Cc ** -0.5)
    b2 = _rand((C2,), torch.float32, 25)
    out = torch.empty(B, C2, T, H, W, device="cuda", dtype=torch.float32)
    ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, head_w2=w2, head_b2=b2, head_exp=True, out_f32=out)
    torch.cuda.synchronize()
    hid = _conv_ref(x, w, (3, 3, 3), b).clamp_min(0)
    ref = torch.exp(hid @ w2.t() + b2).permute(0, 4, 1, 2, 3)
    _close(out, ref, 5e-5)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("stride", [(2, 4, 4), (2, 2, 2), (2, 1, 1), (1, 2, 2)])
def test_conv_transpose(dtype, stride):
    ops = _ops()
    B, T, H, W, Cin, Cout = 1, 4, 8, 8, 256, 64
    sT, sH, sW = stride
    x = _rand((B, T, H, W, Cin), dtype, 26)
    wt = _rand((Cin, Cout, sT, sH, sW), dtype, 27, Cin ** -0.5)  # torch ConvTranspose3d layout
    b = _rand((Cout,), torch.float32, 28)
    wk = wt.permute(2, 3, 4, 1, 0).reshape(sT * sH * sW * Cout, Cin).contiguous()
    bk = b.repeat(sT * sH * sW).contiguous()
    out = torch.empty(B, T * sT, H * sH, W * sW, Cout, device="cuda", dtype=dtype)
    ops.conv_transpose3d(x, wk, bk, stride, out)
    torch.cuda.synchronize()
    ref = F.conv_transpose3d(x.float().permute(0, 4, 1, 2, 3), wt.float(), b, stride=stride).permute(0, 2, 3, 4, 1)
    _close(out, ref, 2 ** -8 if dtype == torch.bfloat16 else 2 ** -10)


@pytest.mark.parametrize("dtype", DT)
def test_splitk_conv_and_linear(dtype):
    """Few output tiles + long K (low-resolution DPT pyramid levels): the split-K path (fp32 atomics into the per-stream
    workspace + finalize kernel) must match the fused single-pass epilogue, and leave its workspace zeroed."""
    ops = _ops()
    from l4p_b200 import ops as O
    B, T, H, W, Cin, Cout = 1, 4, 8, 8, 1024, 256
    x = _rand((B, T, H, W, Cin), dtype, 50)
    w = _rand((Cout, 27 * Cin), dtype, 51, (27 * Cin) ** -0.5)
    b = _rand((Cout,), torch.float32, 52)
    r1 = _rand((B, T, H, W, Cout), dtype, 53)
    o = torch.empty(B, T, H, W, Cout, device="cuda", dtype=dtype)
    orl = torch.empty_like(o)
    ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, res_16=r1, out_16=o, out_16_relu=orl)
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, (3, 3, 3), b) + r1.float()
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -10
    _close(o, ref, tol)
    _close(orl, ref.clamp_min(0), tol)
    assert all(float(ws.abs().max()) == 0.0 for ws in O._SPLITK_WS.values()), "split-K workspace not re-zeroed"
    # same problem with split-K disabled: results agree to fp32 summation-order noise
    O.SPLITK = False
    try:
        o2 = torch.empty_like(o)
        ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, res_16=r1, out_16=o2)
    finally:
        O.SPLITK = True
    torch.cuda.synchronize()
    _close(o, o2.float(), 2 * tol)  # two independently rounded 16-bit results may differ by one ulp
    M, N, K = 256, 1024, 27648
    a = _rand((M, K), dtype, 54)
    wl = _rand((N, K), dtype, 55, K ** -0.5)
    y = torch.empty(M, N, device="cuda", dtype=dtype)
    ops.linear(a, wl, bias=_rand((N,), torch.float32, 56), act=2, out_16=y)
    torch.cuda.synchronize()
    refl = (a.float() @ wl.float().t() + _rand((N,), torch.float32, 56)).clamp_min(0)
    _close(y, refl, tol)


@pytest.mark.parametrize("dtype", DT)
def test_conv_transpose_hyper(dtype):
    """Mask-decoder tail (sam/mask_decoder.py:62-66,137-139): ConvT(k==s) -> GELU -> per-query hyper-network dot."""
    ops = _ops()
    from l4p_b200 import lib
    G, T, H, W, Cin, Cout, C2 = 3, 2, 8, 8, 352, 176, 3
    stride = (1, 2, 2)
    x = _rand((G, T, H, W, Cin), dtype, 40)
    wt = _rand((Cin, Cout, *stride), dtype, 41, Cin ** -0.5)
    b = _rand((Cout,), torch.float32, 42)
    hyper = _rand((G, C2, Cout), torch.float32, 43, Cout ** -0.5)
    wk = wt.permute(2, 3, 4, 1, 0).reshape(-1, Cin).contiguous()
    bk = b.repeat(stride[0] * stride[1] * stride[2]).contiguous()
    out = torch.empty(G, C2, T * stride[0], H * stride[1], W * stride[2], device="cuda", dtype=torch.float32)
    ops.conv_transpose3d_hyper(x, wk, bk, stride, hyper, out, act=lib.ACT_GELU)
    torch.cuda.synchronize()
    up = F.gelu(F.conv_transpose3d(x.float().permute(0, 4, 1, 2, 3), wt.float(), b, stride=stride))  # [G,Cout,T',H',W']
    ref = torch.einsum("gcthw,gkc->gkthw", up, hyper)
    _close(out, ref, 2e-4)


@pytest.mark.parametrize("dtype", DT)
def test_layernorm(dtype):
    ops = _ops()
    x = _rand((2048, 1408), torch.float32, 29, 3.0) + 0.5
    g = _rand((1408,), torch.float32, 30) + 1.0
    b = _rand((1408,), torch.float32, 31)
    o16 = torch.empty(2048, 1408, device="cuda", dtype=dtype)
    o32 = torch.empty(2048, 1408, device="cuda", dtype=torch.float32)
    ops.layernorm(x, g, b, 1e-6, out16=o16, out32=o32)
    torch.cuda.synchronize()
    ref = F.layer_norm(x, (1408,), g, b, 1e-6)
    _close(o32, ref, 2e-6)
    _close(o16, ref, 2 ** -8 if dtype == torch.bfloat16 else 2 ** -10)


def _attn_ref(q, k, v, scale):
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    return torch.softmax(s, dim=-1) @ v.float()


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,H,N,peaky", [(1, 2, 256, 1.0), (1, 16, 2048, 1.0), (2, 3, 512, 6.0), (1, 2, 1024, 30.0)])
def test_attention(dtype, B, H, N, peaky):
    """Fused attention vs fp32 softmax attention on the same rounded q,k,v. `peaky` scales q so the row max
    moves by many log2 units across key blocks (exercises the lazy TMEM rescale)."""
    ops = _ops()
    d, dp = 88, 96
    q = torch.zeros(B, H, N, dp, device="cuda", dtype=dtype)
    k = torch.zeros_like(q)
    vt = torch.zeros(B, H, dp, N, device="cuda", dtype=dtype)
    q[..., :d] = _rand((B, H, N, d), dtype, 40, peaky)
    k[..., :d] = _rand((B, H, N, d), dtype, 41)
    v = _rand((B, H, N, d), dtype, 42)
    vt[:, :, :d] = v.transpose(-1, -2)
    out = torch.empty(B * N, H * d, device="cuda", dtype=dtype)
    scale = d ** -0.5
    ops.attention(q, k, vt, out, d, scale)
    torch.cuda.synchronize()
    ref = _attn_ref(q[..., :d], k[..., :d], v, scale)  # [B,H,N,d]
    ref = ref.permute(0, 2, 1, 3).reshape(B * N, H * d)
    # P is rounded to the operand type before the PV contraction: tolerance = a few operand ulps of |v|max
    _close(out, ref, 2 ** -6 if dtype == torch.bfloat16 else 2 ** -9)
    rel_l2 = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel_l2 < (6e-3 if dtype == torch.bfloat16 else 1e-3), rel_l2


def test_attention_split_variant():
    """The opt-in four-warpgroup softmax variant (L4P_ATT_SPLIT=1: two half-row warpgroups per query tile, partial row maxima
    exchanged through shared memory) against fp32 attention, incl. a peaky case that triggers the split TMEM rescale. The
    selector is read once per process, so the variant runs in a subprocess (tools/att_ab.py prints rel-L2 per case)."""
    import os
    import re
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "tools" / "att_ab.py"), "L4P_ATT_SPLIT", "1"], cwd=root, capture_output=True,
                       text=True, timeout=280, env=dict(os.environ))
    assert r.returncode == 0 and "exit 0" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    rels = [(m.group(1), float(m.group(2))) for m in re.finditer(r"(float16|bfloat16) peaky=[\d.]+: rel-L2 ([\d.e+-]+)", r.stdout)]
    assert len(rels) == 5, r.stdout
    for dt, rel in rels:
        assert rel < (6e-3 if dt == "bfloat16" else 1e-3), (dt, rel)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape,out,align", [
    ((1, 16, 128, 128, 128), (16, 224, 224), True),     # DPT head1 -> head2 resize (dpt_head.py:81-83)
    ((2, 16, 32, 32, 256), (16, 64, 64), True),          # RefineNet x2 (dpt_block.py:231-236), T kept
    ((1, 4, 8, 8, 256), (8, 16, 16), True),              # low-resolution level: T scaled too (all eight corners)
    ((1, 3, 9, 10, 64), (5, 30, 23), False),             # ragged sizes (rows not a multiple of the 4 rows per thread), half-pixel
])
def test_upsample3d(dtype, shape, out, align):
    """Channels-last trilinear resampling (K9) vs F.interpolate on the same rounded input, incl. the fused ReLU copy."""
    import torch.nn.functional as F

    ops = _ops()
    B, T, H, W, C = shape
    x = _rand(shape, dtype, 77)
    y = torch.empty(B, *out, C, device="cuda", dtype=dtype)
    yr = torch.empty_like(y)
    ops.upsample3d(x, out, align_corners=align, y=y, y_relu=yr)
    torch.cuda.synchronize()
    ref = F.interpolate(x.float().permute(0, 4, 1, 2, 3), size=out, mode="trilinear", align_corners=align).permute(0, 2, 3, 4, 1)
    tol = 2 ** -7 if dtype == torch.bfloat16 else 2 ** -10
    _close(y, ref, tol)
    _close(yr, ref.clamp_min(0), tol)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M", [6144, 12288, 2048 * 9])
def test_linear_16bit_residual_stream(dtype, M):
    """The epilogue combinations of the track head's 16-bit token stream (sparse_heads.py step (4)) at its shapes
    (N = 1408, K = 704, M = queries x 2048): fp32 broadcast residual (row % P) -> 16-bit store, 16-bit residual -> 16-bit
    store, and the 16-bit LayerNorm over 1408 columns that follows."""
    ops = _ops()
    N, K, P = 1408, 704, 2048
    a = _rand((M, K), dtype, 11)
    w = _rand((N, K), dtype, 12, K ** -0.5)
    b = _rand((N,), torch.float32, 13)
    table = _rand((P, N), torch.float32, 14)
    out = torch.empty(M, N, device="cuda", dtype=dtype)
    ops.linear(a, w, bias=b, res_f32=table, res_row_mod=P, out_16=out)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + b + table.repeat(M // P, 1)
    tol = 2 ** -7 if dtype == torch.bfloat16 else 2 ** -10
    _close(out, ref, tol)
    res16 = _rand((M, N), dtype, 15)
    out2 = torch.empty_like(out)
    ops.linear(a, w, bias=b, res_16=res16, out_16=out2)
    torch.cuda.synchronize()
    _close(out2, a.float() @ w.float().t() + b + res16.float(), tol)
    g = 1.0 + 0.1 * _rand((N,), torch.float32, 16)
    be = 0.02 * _rand((N,), torch.float32, 17)
    y = torch.empty_like(out2)
    ops.layernorm16(out2, g, be, 1e-5, y)
    torch.cuda.synchronize()
    _close(y, F.layer_norm(out2.float(), (N,), g, be, 1e-5), tol)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,gelu", [(4101, False), (4096 + 16 * 7 + 3, True)])
def test_layernorm16_long_rows_ragged(dtype, rows, gelu):
    """The 4-rows-per-warp 16-bit LayerNorm (1408-channel token stream, rows >= 4096) with a row count that is not a multiple
    of the 16 rows of a block: the tail rows are written, nothing beyond them is."""
    ops = _ops()
    N = 1408
    x = _rand((rows, N), dtype, 21, 1.7) + 0.3
    g = 1.0 + 0.1 * _rand((N,), torch.float32, 22)
    be = 0.02 * _rand((N,), torch.float32, 23)
    y = torch.full((rows + 16, N), 7.0, device="cuda", dtype=dtype)
    ops.layernorm16(x, g, be, 1e-5, y[:rows], gelu=gelu)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (N,), g, be, 1e-5)
    if gelu:
        ref = F.gelu(ref)
    _close(y[:rows], ref, 2 ** -7 if dtype == torch.bfloat16 else 2 ** -10)
    assert bool((y[rows:] == 7.0).all())


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(1, 4, 64, 64, 128, 128), (1, 3, 20, 32, 64, 128), (2, 2, 12, 96, 256, 64),
                                   (1, 1, 8, 32, 64, 16)])
def test_conv3d_line_halo_pair(dtype, shape):
    """Narrow 3x3x3 convolutions through the 2-CTA kernel's line-halo stages (one 6-line A box per (dt, dw, channel block)
    serves the three row taps): whole volumes, an odd tile count (padding block of the last pair), several channel blocks,
    a single frame (the dt = +-1 boxes lie entirely outside the volume: TMA zero fill)."""
    ops = _ops()
    B, T, H, W, Cin, Cout = shape
    x = _rand((B, T, H, W, Cin), dtype, 31)
    w = _rand((Cout, 27 * Cin), dtype, 32, (27 * Cin) ** -0.5)
    b = _rand((Cout,), torch.float32, 33)
    out = torch.empty(B, T, H, W, Cout, device="cuda", dtype=torch.float32)
    ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, out_f32=out, cta_pair=1)
    torch.cuda.synchronize()
    _close(out, _conv_ref(x, w, (3, 3, 3), b), 3e-5)


@pytest.mark.parametrize("dtype", DT)
def test_conv3d_head1x1_line_halo_pair(dtype):
    """The DPT head convolution (3x3x3 128 -> 128 + ReLU + 1x1x1, fused-dot epilogue) at a size that takes the 2-CTA kernel."""
    ops = _ops()
    B, T, H, W, Cc, C2 = 1, 2, 56, 64, 128, 2
    x = _rand((B, T, H, W, Cc), dtype, 34)
    w = _rand((Cc, 27 * Cc), dtype, 35, (27 * Cc) ** -0.5)
    b = _rand((Cc,), torch.float32, 36)
    w2 = _rand((C2, Cc), torch.float32, 37, Cc ** -0.5)
    b2 = _rand((C2,), torch.float32, 38)
    out = torch.empty(B, C2, T, H, W, device="cuda", dtype=torch.float32)
    ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, head_w2=w2, head_b2=b2, head_exp=False, out_f32=out, cta_pair=1)
    torch.cuda.synchronize()
    hid = _conv_ref(x, w, (3, 3, 3), b).clamp_min(0)
    ref = (hid @ w2.t() + b2).permute(0, 4, 1, 2, 3)
    _close(out, ref, 5e-5)


@pytest.mark.parametrize("dtype", DT)
def test_conv3d_line_halo_wide(dtype):
    """Line-halo stages with 256-wide tiles (two 72 KiB stages in the ring): the 64^2 RefineNet convolution shape class."""
    ops = _ops()
    B, T, H, W, Cin, Cout = 1, 3, 32, 64, 128, 256
    x = _rand((B, T, H, W, Cin), dtype, 41)
    w = _rand((Cout, 27 * Cin), dtype, 42, (27 * Cin) ** -0.5)
    b = _rand((Cout,), torch.float32, 43)
    r1 = _rand((B, T, H, W, Cout), dtype, 44)
    o = torch.empty(B, T, H, W, Cout, device="cuda", dtype=dtype)
    ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, res_16=r1, out_16=o, cta_pair=1)
    torch.cuda.synchronize()
    _close(o, _conv_ref(x, w, (3, 3, 3), b) + r1.float(), 2 ** -8 if dtype == torch.bfloat16 else 2 ** -10)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape,pair", [((2, 2, 16, 32, 64, 64), 0), ((1, 4, 64, 64, 128, 128), 1), ((2, 1, 8, 8, 256, 256), 0),
                                        ((1, 3, 20, 32, 64, 32), -1)])
def test_conv3d_grouped_weights(dtype, shape, pair):
    """Grouped convolution (several heads' identical layers in one launch): G = 3 blocks of the batch axis, each with its own
    weight block, bias and residual; 1-CTA, 2-CTA (line-halo) and split-K paths."""
    ops = _ops()
    G = 3
    Bg, T, H, W, Cin, Cout = shape
    x = _rand((G * Bg, T, H, W, Cin), dtype, 51)
    w = _rand((G * Cout, 27 * Cin), dtype, 52, (27 * Cin) ** -0.5)
    b = _rand((G * Cout,), torch.float32, 53)
    r = _rand((G * Bg, T, H, W, Cout), dtype, 54)
    o = torch.empty(G * Bg, T, H, W, Cout, device="cuda", dtype=dtype)
    orl = torch.empty_like(o)
    ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, res_16=r, out_16=o, out_16_relu=orl, groups=G, cta_pair=pair)
    torch.cuda.synchronize()
    tol = 2 ** -7 if dtype == torch.bfloat16 else 2 ** -10
    for g in range(G):
        sl = slice(g * Bg, (g + 1) * Bg)
        ref = _conv_ref(x[sl], w[g * Cout:(g + 1) * Cout], (3, 3, 3), b[g * Cout:(g + 1) * Cout]) + r[sl].float()
        _close(o[sl], ref, tol)
        _close(orl[sl], ref.clamp_min(0), tol)
    # 1x1x1 grouped conv (the fusion blocks' out_conv)
    w1 = _rand((G * Cout, Cin), dtype, 55, Cin ** -0.5)
    o1 = torch.empty(G * Bg, T, H, W, Cout, device="cuda", dtype=dtype)
    ops.conv3d(x, w1, ksize=(1, 1, 1), bias=b, out_16=o1, groups=G)
    torch.cuda.synchronize()
    for g in range(G):
        sl = slice(g * Bg, (g + 1) * Bg)
        ref = x[sl].float() @ w1[g * Cout:(g + 1) * Cout].float().t() + b[g * Cout:(g + 1) * Cout]
        _close(o1[sl], ref, tol)
