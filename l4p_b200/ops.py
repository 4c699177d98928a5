"""Thin torch-facing wrappers over the C ABI: torch only supplies device memory and the current stream.

Every function validates its tensors (CUDA, contiguous, dtype) and raises on any library error;
nothing here computes on the CPU or falls back to ATen.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import lib as _l

_initialised = set()

# --- instrumentation used by bench.py (never changes results) ------------------------------------------
LAUNCHES = 0            # kernels launched through this module since import (bench.py reports the delta)
ATTN_EVENTS = None      # set to a list to record (start, end) CUDA events around every attention launch
GEMM_PAIR = int(__import__("os").environ.get("L4P_GEMM_PAIR", "0"))  # 0 auto, 1 force 2-CTA tiles, -1 never (tuning/tests)


def _dev_init(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise _l.L4PError("l4p_b200 ops need CUDA tensors (no CPU fallback exists for this path)")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    # One process per GPU: kernels are launched on the CURRENT device's current stream and the
    # library's per-function attributes / SM count are per process, so operands on another device are an error, not a
    # silent launch on the wrong stream.
    cur = torch.cuda.current_device()
    if idx != cur:
        raise _l.L4PError(f"tensor on cuda:{idx} but the current device is cuda:{cur}: l4p_b200 runs one process per GPU "
                          f"(call torch.cuda.set_device first)")
    if idx not in _initialised:
        _l.check(_l.load().l4p_init(idx, 1), "l4p_init")
        _initialised.add(idx)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _is16(t: torch.Tensor) -> bool:
    return t.dtype in (torch.float16, torch.bfloat16)


def _chk(t: Optional[torch.Tensor], name: str, dtype=None, sixteen=False) -> None:
    if t is None:
        return
    if not t.is_cuda or not t.is_contiguous():
        raise _l.L4PError(f"{name}: expected a contiguous CUDA tensor")
    if dtype is not None and t.dtype != dtype:
        raise _l.L4PError(f"{name}: expected {dtype}, got {t.dtype}")
    if sixteen and not _is16(t):
        raise _l.L4PError(f"{name}: expected fp16/bf16, got {t.dtype}")


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out16: Optional[torch.Tensor] = None, out32: Optional[torch.Tensor] = None) -> None:
    """K2. x fp32 [..., C] -> out16 (fp16/bf16) and/or out32."""
    _dev_init(x)
    _chk(x, "x", torch.float32); _chk(gamma, "gamma", torch.float32); _chk(beta, "beta", torch.float32)
    _chk(out16, "out16", sixteen=True); _chk(out32, "out32", torch.float32)
    cols = x.shape[-1]
    rows = x.numel() // cols
    bf16 = 1 if (out16 is not None and out16.dtype == torch.bfloat16) else 0
    _l.check(_l.load().l4p_layernorm(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(out16), _ptr(out32), rows, cols,
                                     float(eps), bf16, _stream()), "l4p_layernorm")
    _count()


# zero-filled fp32 split-K workspaces, one per (device, stream): kernels on different streams may run concurrently.
# l4p_gemm's finalize kernel re-zeroes what it used, so a buffer stays valid for the next call on the same stream.
_SPLITK_WS: Dict[Tuple[int, int], torch.Tensor] = {}
SPLITK_WS_BYTES = 8 << 20   # covers M*N*4 of every low-resolution pyramid level (largest: 4096 x 256 fp32 = 4 MiB)
SPLITK = True




def _splitk_ws(dev: torch.device) -> torch.Tensor:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream() or 0)
    ws = _SPLITK_WS.get(key)
    if ws is None:
        ws = torch.zeros(SPLITK_WS_BYTES // 4, device=dev, dtype=torch.float32)
        _SPLITK_WS[key] = ws
    return ws


def _base_desc(a: torch.Tensor, w: torch.Tensor) -> _l.GemmDesc:
    _dev_init(a)
    _chk(a, "a", sixteen=True); _chk(w, "w", sixteen=True)
    if a.dtype != w.dtype:
        raise _l.L4PError(f"operand dtypes differ: {a.dtype} vs {w.dtype}")
    d = _l.GemmDesc()
    d.a = a.data_ptr(); d.w = w.data_ptr()
    d.cta_pair = GEMM_PAIR
    d.bf16 = 1 if a.dtype == torch.bfloat16 else 0
    if SPLITK:
        ws = _splitk_ws(a.device)
        d.splitk_ws, d.splitk_ws_bytes, d.split_k = ws.data_ptr(), ws.numel() * 4, 0
    else:
        d.split_k = 1
    return d


def _epilogue(d: _l.GemmDesc, *, bias=None, act=_l.ACT_NONE, res_f32=None, res_16=None, res2_16=None,
              out_f32=None, out_16=None, out_16_relu=None) -> None:
    _chk(bias, "bias", torch.float32); _chk(res_f32, "res_f32", torch.float32)
    _chk(res_16, "res_16", sixteen=True); _chk(res2_16, "res2_16", sixteen=True)
    _chk(out_f32, "out_f32", torch.float32); _chk(out_16, "out_16", sixteen=True)
    _chk(out_16_relu, "out_16_relu", sixteen=True)
    d.bias = _ptr(bias); d.act = act
    d.res_f32 = _ptr(res_f32); d.res_16 = _ptr(res_16); d.res2_16 = _ptr(res2_16)
    d.out_f32 = _ptr(out_f32); d.out_16 = _ptr(out_16); d.out_16_relu = _ptr(out_16_relu)
    d.ld_res = d.N; d.ld_out = d.N


def linear(a: torch.Tensor, w: torch.Tensor, *, bias=None, act=_l.ACT_NONE, res_f32=None, res_16=None,
           out_f32=None, out_16=None, out_16_relu=None, block_n: int = 0, res_row_mod: int = 0,
           cta_pair: int = 0, prof: Optional[torch.Tensor] = None, group_rows: int = 0, n_out: int = 0) -> None:
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias) (+ residual). Replaces F.linear/addmm call sites.

    group_rows > 0: grouped weights. Rows [g*group_rows, (g+1)*group_rows) of `a` multiply their own block
    w[g*Nw:(g+1)*Nw] of a [groups*Nw, K] weight stack (Nw = w.shape[0] // groups); the output has n_out (default Nw) columns."""
    d = _base_desc(a, w)
    K = a.shape[-1]
    M = a.numel() // K
    N = w.shape[0]
    if w.shape[1] != K:
        raise _l.L4PError(f"linear: K mismatch {w.shape} vs {a.shape}")
    if group_rows > 0:
        groups = -(-M // group_rows)
        if M % group_rows or N % groups:
            raise _l.L4PError(f"linear(grouped): M={M} / group_rows={group_rows}, w rows {N} / groups {groups}")
        d.grp_a_rows, d.grp_b_rows = group_rows, N // groups
        d.m_stride = group_rows if group_rows < 128 else 0
        N = n_out or N // groups
        d.split_k = 1
    d.M, d.N, d.K = M, N, K
    d.lda, d.ldw = K, K
    d.a_mode = _l.A_MATRIX
    d.store_mode = _l.STORE_ROWMAJOR
    d.block_n = block_n
    d.res_row_mod = res_row_mod
    d.cta_pair = cta_pair
    d.prof = _ptr(prof)
    _epilogue(d, bias=bias, act=act, res_f32=res_f32, res_16=res_16, out_f32=out_f32, out_16=out_16,
              out_16_relu=out_16_relu)
    _l.check(_l.load().l4p_gemm(C.byref(d), _stream()), "l4p_gemm(linear)")
    _count()


def linear_qkv(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, q: torch.Tensor, k: torch.Tensor,
               vt: torch.Tensor, heads: int, head_dim: int, tokens: int, block_n: int = 0) -> None:
    """K3: qkv projection with the attention layout fused into the epilogue.

    a [B*tokens, C]; w [3*heads*head_dim, C]; q,k [B,heads,tokens,dpad]; vt [B,heads,dpad,tokens]
    (pad lanes are never written: allocate them zero-filled once)."""
    d = _base_desc(a, w)
    K = a.shape[-1]
    d.M, d.N, d.K = a.numel() // K, w.shape[0], K
    d.lda, d.ldw = K, K
    d.a_mode = _l.A_MATRIX
    d.store_mode = _l.STORE_QKV
    _chk(bias, "bias", torch.float32)
    for n, t in (("q", q), ("k", k), ("vt", vt)):
        _chk(t, n, sixteen=True)
        if t.dtype != a.dtype:
            raise _l.L4PError(f"{n}: dtype {t.dtype} != operand dtype {a.dtype}")
    d.bias = _ptr(bias)
    d.q, d.k, d.vt = q.data_ptr(), k.data_ptr(), vt.data_ptr()
    d.heads, d.head_dim, d.head_dim_pad, d.tokens = heads, head_dim, q.shape[-1], tokens
    d.block_n = block_n
    _l.check(_l.load().l4p_gemm(C.byref(d), _stream()), "l4p_gemm(qkv)")
    _count()


def pick_box(T: int, H: int, W: int) -> Tuple[int, int, int]:
    """(bT,bH,bW) voxel box with 128 voxels that tiles [T,H,W] with the least padding."""
    best = None
    for bw in (128, 64, 32, 16, 8, 4, 2, 1):
        for bh in (128, 64, 32, 16, 8, 4, 2, 1):
            if bw * bh > 128 or 128 % (bw * bh):
                continue
            bt = 128 // (bw * bh)
            waste = (-(-T // bt) * bt) * (-(-H // bh) * bh) * (-(-W // bw) * bw)
            key = (waste, -bw, -bh)
            if best is None or key < best[0]:
                best = (key, (bt, bh, bw))
    return best[1]


def conv3d(x: torch.Tensor, w: torch.Tensor, *, ksize: Tuple[int, int, int], bias=None, act=_l.ACT_NONE,
           res_16=None, res2_16=None, out_16=None, out_16_relu=None, out_f32=None,
           head_w2=None, head_b2=None, head_exp=False, block_n: int = 0, cta_pair: int = 0,
           prof: Optional[torch.Tensor] = None, groups: int = 1) -> None:
    """K8: stride-1 'same' Conv3d as implicit GEMM.

    x channels-last [B,T,H,W,Cin] (16-bit); w [Cout, kT*kH*kW*Cin] with the K axis ordered (kt,kh,kw,cin).
    With head_w2/head_b2 the epilogue is ReLU -> 1x1x1 conv (+exp) -> out_f32 [B,C2,T,H,W].
    groups > 1: the batch axis holds `groups` equal blocks, block g convolves with w[g*Cout:(g+1)*Cout] (w [groups*Cout, K],
    bias [groups*Cout]): the identical layers of several heads in one launch."""
    d = _base_desc(x, w)
    B, T, H, W, Cin = x.shape
    kT, kH, kW = ksize
    Cout = w.shape[0]
    if groups > 1:
        if B % groups or Cout % groups or head_w2 is not None:
            raise _l.L4PError(f"conv3d(groups={groups}): batch {B} / weight rows {Cout} not divisible, or fused head epilogue")
        Cout //= groups
        d.conv_grp_b = B // groups
    if w.shape[1] != kT * kH * kW * Cin:
        raise _l.L4PError(f"conv3d: weight {tuple(w.shape)} vs taps*Cin={kT * kH * kW * Cin}")
    d.M, d.N, d.K = B * T * H * W, Cout, kT * kH * kW * Cin
    d.lda, d.ldw = Cin, w.shape[1]
    d.a_mode = _l.A_CONV3D
    d.cB, d.cT, d.cH, d.cW, d.cCin = B, T, H, W, Cin
    d.kT, d.kH, d.kW = kT, kH, kW
    d.bT, d.bH, d.bW = pick_box(T, H, W)
    if kH == 3 and kW == 3 and H % 4 == 0 and W % 32 == 0:
        # a 4-line box lets the 2-CTA kernel's line-halo stages serve the three row taps from one 6-line A box (l4p_gemm,
        # gemm2_kernel); the wide single- or two-line boxes pick_box prefers at W = 128 / 64 have little or no row reuse
        d.bT, d.bH, d.bW = 1, 4, 32
    d.block_n = block_n
    d.cta_pair = cta_pair
    d.prof = _ptr(prof)
    if head_w2 is not None:
        _chk(head_w2, "head_w2", torch.float32); _chk(head_b2, "head_b2", torch.float32)
        _chk(out_f32, "out_f32", torch.float32); _chk(bias, "bias", torch.float32)
        d.store_mode = _l.STORE_HEAD1X1
        d.bias = _ptr(bias); d.act = _l.ACT_RELU
        d.w2, d.b2, d.c2, d.exp_out = head_w2.data_ptr(), head_b2.data_ptr(), head_w2.shape[0], int(head_exp)
        d.out_f32 = out_f32.data_ptr()
    else:
        d.store_mode = _l.STORE_ROWMAJOR
        _epilogue(d, bias=bias, act=act, res_16=res_16, res2_16=res2_16, out_f32=out_f32, out_16=out_16,
                  out_16_relu=out_16_relu)
    _l.check(_l.load().l4p_gemm(C.byref(d), _stream()), "l4p_gemm(conv3d)")
    _count()


def conv_transpose3d(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, stride: Tuple[int, int, int],
                     out_16: torch.Tensor, block_n: int = 0) -> None:
    """K7: ConvTranspose3d with kernel == stride as GEMM + pixel-shuffle store.

    x channels-last [B,T,H,W,Cin]; w [sT*sH*sW*Cout, Cin] with rows ordered (kt,kh,kw,co);
    bias [sT*sH*sW*Cout] (the per-channel bias tiled over taps); out_16 [B,T*sT,H*sH,W*sW,Cout]."""
    d = _base_desc(x, w)
    B, T, H, W, Cin = x.shape
    sT, sH, sW = stride
    Cout = w.shape[0] // (sT * sH * sW)
    d.M, d.N, d.K = B * T * H * W, w.shape[0], Cin
    d.lda, d.ldw = Cin, Cin
    d.a_mode = _l.A_MATRIX
    d.store_mode = _l.STORE_CONVT
    d.cB, d.cT, d.cH, d.cW = B, T, H, W
    d.sT, d.sH, d.sW, d.ctCout = sT, sH, sW, Cout
    _chk(bias, "bias", torch.float32); _chk(out_16, "out_16", sixteen=True)
    d.bias = _ptr(bias); d.out_16 = out_16.data_ptr()
    d.block_n = block_n
    _l.check(_l.load().l4p_gemm(C.byref(d), _stream()), "l4p_gemm(convT)")
    _count()


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, out: torch.Tensor, head_dim: int,
              scale: float, prof: Optional[torch.Tensor] = None) -> None:
    """K4: fused softmax(q k^T scale) v. q,k [B,H,N,dpad]; vt [B,H,dpad,N]; out [B*N, H*head_dim] 16-bit."""
    _dev_init(q)
    for n, t in (("q", q), ("k", k), ("vt", vt), ("out", out)):
        _chk(t, n, sixteen=True)
        if t.dtype != q.dtype:
            raise _l.L4PError(f"attention: {n} dtype {t.dtype} != {q.dtype}")
    B, H, N, dpad = q.shape
    if tuple(k.shape) != (B, H, N, dpad) or tuple(vt.shape) != (B, H, dpad, N):
        raise _l.L4PError(f"attention: shapes q{tuple(q.shape)} k{tuple(k.shape)} vt{tuple(vt.shape)}")
    if out.numel() != B * N * H * head_dim:
        raise _l.L4PError(f"attention: out has {out.numel()} elements, expected {B * N * H * head_dim}")
    ev = None
    if ATTN_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    _l.check(_l.load().l4p_attention(q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr(), B, H, N, head_dim,
                                     dpad, float(scale), 1 if q.dtype == torch.bfloat16 else 0, _stream(), _ptr(prof)),
             "l4p_attention")
    if ev is not None:
        ev[1].record()
        ATTN_EVENTS.append((ev[0], ev[1], B * H * N * N * head_dim * 4))
    _count()


def patchify(rgb: torch.Tensor, out16: torch.Tensor, tubelet: Tuple[int, int, int]) -> None:
    """K1 gather: rgb fp32 [B,C,T,H,W] -> out16 [B*tokens, C*pt*ph*pw]."""
    _dev_init(rgb)
    _chk(rgb, "rgb", torch.float32); _chk(out16, "out16", sixteen=True)
    B, Cc, T, H, W = rgb.shape
    pt, ph, pw = tubelet
    if out16.numel() != rgb.numel():
        raise _l.L4PError("patchify: output size mismatch")
    _l.check(_l.load().l4p_patchify(rgb.data_ptr(), out16.data_ptr(), B, Cc, T, H, W, pt, ph, pw,
                                    1 if out16.dtype == torch.bfloat16 else 0, _stream()), "l4p_patchify")
    _count()


def cast16(x: torch.Tensor, y16: torch.Tensor, add: Optional[torch.Tensor] = None) -> None:
    """y16 = round16(x (+ add)); x, add fp32 of the same size."""
    _dev_init(x)
    _chk(x, "x", torch.float32); _chk(y16, "y16", sixteen=True); _chk(add, "add", torch.float32)
    if x.numel() != y16.numel() or (add is not None and add.numel() != x.numel()):
        raise _l.L4PError("cast16: size mismatch")
    _l.check(_l.load().l4p_cast16_add(x.data_ptr(), _ptr(add), y16.data_ptr(), x.numel(),
                                      1 if y16.dtype == torch.bfloat16 else 0, _stream()), "l4p_cast16")
    _count()


def upsample3d(x: torch.Tensor, out_size: Tuple[int, int, int], *, align_corners: bool, y: Optional[torch.Tensor] = None,
               y_relu: Optional[torch.Tensor] = None) -> None:
    """K9: trilinear resampling of channels-last [B,T,H,W,C] 16-bit to out_size."""
    _dev_init(x)
    _chk(x, "x", sixteen=True); _chk(y, "y", sixteen=True); _chk(y_relu, "y_relu", sixteen=True)
    B, Ti, Hi, Wi, Cc = x.shape
    To, Ho, Wo = out_size
    for t in (y, y_relu):
        if t is not None and (t.numel() != B * To * Ho * Wo * Cc or t.dtype != x.dtype):
            raise _l.L4PError("upsample3d: output size/dtype mismatch")
    _l.check(_l.load().l4p_upsample3d(x.data_ptr(), _ptr(y), _ptr(y_relu), B, Ti, Hi, Wi, To, Ho, Wo, Cc,
                                      int(align_corners), 1 if x.dtype == torch.bfloat16 else 0, _stream()),
             "l4p_upsample3d")
    _count()


def im2col3(x: torch.Tensor, out: torch.Tensor, stride: Tuple[int, int, int]) -> None:
    """3x3x3/pad-1 strided gather of channels-last x [B,T,H,W,C] into out [B*To*Ho*Wo, 27*C]."""
    _dev_init(x)
    _chk(x, "x", sixteen=True); _chk(out, "out", sixteen=True)
    B, T, H, W, Cc = x.shape
    _l.check(_l.load().l4p_im2col3(x.data_ptr(), out.data_ptr(), B, T, H, W, Cc, *stride, _stream()), "l4p_im2col3")
    _count()


def conv_transpose3d_hyper(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, stride: Tuple[int, int, int],
                           hyper: torch.Tensor, out_f32: torch.Tensor, act=_l.ACT_GELU,
                           prof: Optional[torch.Tensor] = None) -> None:
    """K14: last ConvTranspose3d of the mask decoder fused with its activation and the hyper-network dot.

    x channels-last [G,T,H,W,Cin]; w [sT*sH*sW*Cout, Cin] rows (kt,kh,kw,co); bias tiled likewise;
    hyper fp32 [G, c2, Cout]; out_f32 [G, c2, T*sT, H*sH, W*sW]."""
    d = _base_desc(x, w)
    G, T, H, W, Cin = x.shape
    sT, sH, sW = stride
    Cout = w.shape[0] // (sT * sH * sW)
    _chk(bias, "bias", torch.float32); _chk(hyper, "hyper", torch.float32); _chk(out_f32, "out_f32", torch.float32)
    if tuple(hyper.shape) != (G, hyper.shape[1], Cout) or out_f32.numel() != G * hyper.shape[1] * T * sT * H * sH * W * sW:
        raise _l.L4PError("conv_transpose3d_hyper: shape mismatch")
    d.M, d.N, d.K = G * T * H * W, w.shape[0], Cin
    d.lda, d.ldw = Cin, Cin
    d.a_mode = _l.A_MATRIX
    d.store_mode = _l.STORE_HYPER
    d.cB, d.cT, d.cH, d.cW = G, T, H, W
    d.sT, d.sH, d.sW, d.ctCout = sT, sH, sW, Cout
    d.bias = _ptr(bias); d.act = act
    d.w2 = hyper.data_ptr(); d.c2 = hyper.shape[1]
    d.rows_per_group = T * H * W
    d.out_f32 = out_f32.data_ptr()
    d.block_n = Cout
    d.prof = _ptr(prof)
    _l.check(_l.load().l4p_gemm(C.byref(d), _stream()), "l4p_gemm(hyper)")
    _count()


def token_attention(q: torch.Tensor, k16: torch.Tensor, v16: torch.Tensor, out: torch.Tensor, heads: int,
                    shared_kv: bool, scale: float) -> None:
    """K13a: q,out fp32 [G,nq,C]; k16,v16 16-bit [G*Nk, C] (or [Nk, C] when shared_kv)."""
    _dev_init(q)
    _chk(q, "q", torch.float32); _chk(out, "out", torch.float32); _chk(k16, "k16", sixteen=True); _chk(v16, "v16", sixteen=True)
    G, nq, Cc = q.shape
    Nk = k16.shape[0] if shared_kv else k16.shape[0] // G
    d = Cc // heads
    _l.check(_l.load().l4p_token_attention(q.data_ptr(), k16.data_ptr(), v16.data_ptr(), out.data_ptr(), G, nq, Nk, heads, d,
                                           0 if shared_kv else Nk, float(scale), 1 if k16.dtype == torch.bfloat16 else 0,
                                           _stream()), "l4p_token_attention")
    _count()


def image_attention(q16: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out16: torch.Tensor, G: int, heads: int,
                    scale: float) -> None:
    """K13b: q16,out16 16-bit [G*Np, C]; k,v fp32 [G,nk,C]."""
    _dev_init(q16)
    _chk(q16, "q16", sixteen=True); _chk(out16, "out16", sixteen=True); _chk(k, "k", torch.float32); _chk(v, "v", torch.float32)
    Np = q16.shape[0] // G
    Cc = q16.shape[1]
    _l.check(_l.load().l4p_image_attention(q16.data_ptr(), k.data_ptr(), v.data_ptr(), out16.data_ptr(), G, Np, k.shape[1],
                                           heads, Cc // heads, float(scale), 1 if q16.dtype == torch.bfloat16 else 0,
                                           _stream()), "l4p_image_attention")
    _count()


def layernorm16(x16: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, y16: torch.Tensor,
                gelu: bool = False) -> None:
    _dev_init(x16)
    _chk(x16, "x16", sixteen=True); _chk(y16, "y16", sixteen=True); _chk(gamma, "gamma", torch.float32); _chk(beta, "beta", torch.float32)
    cols = x16.shape[-1]
    _l.check(_l.load().l4p_layernorm16(x16.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y16.data_ptr(), x16.numel() // cols,
                                       cols, float(eps), int(gelu), 1 if x16.dtype == torch.bfloat16 else 0, _stream()),
             "l4p_layernorm16")
    _count()


def head_expand(q: torch.Tensor, out16: torch.Tensor, G: int, nt: int, heads: int, hd: int, scale: float) -> None:
    """q fp32 [G*nt, heads*hd] -> out16 [G*heads*nt, heads*hd]: row (g,h,t) = scale * q[g,t] on head h's columns, 0 elsewhere."""
    _dev_init(q)
    _chk(q, "q", torch.float32); _chk(out16, "out16", sixteen=True)
    if q.numel() != G * nt * heads * hd or out16.numel() != G * heads * nt * heads * hd:
        raise _l.L4PError(f"head_expand: shapes {tuple(q.shape)} / {tuple(out16.shape)} vs G={G} nt={nt} heads={heads} hd={hd}")
    _l.check(_l.load().l4p_head_expand(q.data_ptr(), out16.data_ptr(), G, nt, heads, hd, float(scale),
                                       1 if out16.dtype == torch.bfloat16 else 0, _stream()), "l4p_head_expand")
    _count()


def head_diag_gather(z: torch.Tensor, out: torch.Tensor, G: int, nt: int, heads: int, hd: int) -> None:
    """z fp32 [G*heads*nt, heads*hd] -> out fp32 [G*nt, heads*hd]: out[(g,t), h*hd+d] = z[(g,h,t), h*hd+d]."""
    _dev_init(z)
    _chk(z, "z", torch.float32); _chk(out, "out", torch.float32)
    if out.numel() != G * nt * heads * hd or z.numel() != G * heads * nt * heads * hd:
        raise _l.L4PError(f"head_diag_gather: shapes {tuple(z.shape)} / {tuple(out.shape)}")
    _l.check(_l.load().l4p_head_diag_gather(z.data_ptr(), out.data_ptr(), G, nt, heads, hd, _stream()), "l4p_head_diag_gather")
    _count()


def row_softmax16(s: torch.Tensor, p16: torch.Tensor) -> None:
    """fp32 scores [rows, n] (already scaled) -> 16-bit probabilities."""
    _dev_init(s)
    _chk(s, "s", torch.float32); _chk(p16, "p16", sixteen=True)
    n = s.shape[-1]
    if p16.shape != s.shape:
        raise _l.L4PError(f"row_softmax16: {tuple(s.shape)} vs {tuple(p16.shape)}")
    _l.check(_l.load().l4p_row_softmax16(s.data_ptr(), p16.data_ptr(), s.numel() // n, n,
                                         1 if p16.dtype == torch.bfloat16 else 0, _stream()), "l4p_row_softmax16")
    _count()


def group_softmax_t16(s: torch.Tensor, p16: torch.Tensor, G: int, heads: int, nt: int) -> None:
    """s fp32 [G*heads*nt, n] (transposed scores) -> p16 [G*n, heads*nt]: softmax over the nt tokens of each head."""
    _dev_init(s)
    _chk(s, "s", torch.float32); _chk(p16, "p16", sixteen=True)
    n = s.shape[-1]
    if s.numel() != G * heads * nt * n or p16.numel() != s.numel():
        raise _l.L4PError(f"group_softmax_t16: shapes {tuple(s.shape)} / {tuple(p16.shape)} vs G={G} heads={heads} nt={nt}")
    _l.check(_l.load().l4p_group_softmax_t16(s.data_ptr(), p16.data_ptr(), G, heads, nt, n,
                                             1 if p16.dtype == torch.bfloat16 else 0, _stream()), "l4p_group_softmax_t16")
    _count()


def token_weighted_sum(p16: torch.Tensor, x16: torch.Tensor, y16: torch.Tensor, G: int, J: int) -> None:
    """y16[g] = p16[g] @ x16[g]: p16 [G*J, n], x16 [G*n, C], y16 [G*J, C] (fp32 accumulation)."""
    _dev_init(p16)
    for nme, t in (("p16", p16), ("x16", x16), ("y16", y16)):
        _chk(t, nme, sixteen=True)
    n, Cc = p16.shape[-1], x16.shape[-1]
    if p16.numel() != G * J * n or x16.numel() != G * n * Cc or y16.numel() != G * J * Cc or not (p16.dtype == x16.dtype == y16.dtype):
        raise _l.L4PError(f"token_weighted_sum: shapes {tuple(p16.shape)} {tuple(x16.shape)} {tuple(y16.shape)} G={G} J={J}")
    _l.check(_l.load().l4p_token_weighted_sum(p16.data_ptr(), x16.data_ptr(), y16.data_ptr(), G, J, n, Cc,
                                              1 if p16.dtype == torch.bfloat16 else 0, _stream()), "l4p_token_weighted_sum")
    _count()


def track_readout(masks: torch.Tensor, image_hw: Tuple[int, int]):
    """K15: masks fp32 [G,nch,T,h,w] -> (traj [G,2,T], vis [G,1,T] | None, depth [G,1,T] | None)."""
    _dev_init(masks)
    _chk(masks, "masks", torch.float32)
    G, nch, T, h, w = masks.shape
    dev = masks.device
    traj = torch.empty(G, 2, T, device=dev, dtype=torch.float32)
    vis = torch.empty(G, 1, T, device=dev, dtype=torch.float32) if nch >= 2 else None
    depth = torch.empty(G, 1, T, device=dev, dtype=torch.float32) if nch >= 3 else None
    _l.check(_l.load().l4p_track_readout(masks.data_ptr(), traj.data_ptr(), _ptr(vis), _ptr(depth), G, nch, T, h, w,
                                         image_hw[0], image_hw[1], _stream()), "l4p_track_readout")
    _count()
    return traj, vis, depth
