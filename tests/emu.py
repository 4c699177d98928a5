"""CPU stand-in for the kernel layer -- TEST INFRASTRUCTURE ONLY (never imported by l4p_b200/).

`install(monkeypatch)` replaces every entry of `l4p_b200.ops` (and the four geometry wrappers that call the C ABI
directly) by a plain-torch function with the SAME signature, the same operand / output dtypes and the same rounding
points (16-bit operands, fp32 accumulation, results rounded once when stored to a 16-bit tensor). The host mirror
(`l4p_b200.models.*`: weight packing, layouts, kernel sequencing, window batching, stitching, the tracker state machine)
then runs unchanged on CPU tensors, so `-m "not gpu"` tests can hold it to the reference goldens in tests/golden/ at tiny
sizes. The definitions below are the per-op references the GPU tests use (tests/test_gemm_gpu.py, tests/test_track_gpu.py);
the geometry solves defer to oracle/ (allowed here: tests may use the oracle as the checker).

The product path is unaffected: without `install`, CPU tensors still raise L4PError (tests/test_host_logic.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU, ACT_RELU, ACT_EXP = 0, 1, 2, 3
CALLS = {}


def _note(name: str) -> None:
    CALLS[name] = CALLS.get(name, 0) + 1


def _act(x: torch.Tensor, act: int) -> torch.Tensor:
    if act == ACT_GELU:
        return F.gelu(x)
    if act == ACT_RELU:
        return x.clamp_min(0)
    if act == ACT_EXP:
        return torch.exp(x)
    return x


def _is16(t: torch.Tensor) -> bool:
    return t.dtype in (torch.float16, torch.bfloat16)


def _store(acc: torch.Tensor, out_f32, out_16, out_16_relu) -> None:
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.numel() == acc.numel()
        out_f32.view(acc.shape).copy_(acc)
    if out_16 is not None:
        assert _is16(out_16) and out_16.numel() == acc.numel()
        out_16.view(acc.shape).copy_(acc.to(out_16.dtype))
    if out_16_relu is not None:
        assert _is16(out_16_relu)
        out_16_relu.view(acc.shape).copy_(acc.clamp_min(0).to(out_16_relu.dtype))


def _epilogue(acc, *, bias=None, act=ACT_NONE, res_f32=None, res_16=None, res2_16=None, res_row_mod=0):
    """acc fp32 [M,N] -> act(acc + bias) + residual(s); residual rows are `row % res_row_mod` of a table when set."""
    M, N = acc.shape
    if bias is not None:
        assert bias.dtype == torch.float32
        acc = acc + bias
    acc = _act(acc, act)
    for r in (res_f32, res_16, res2_16):
        if r is None:
            continue
        r = r.reshape(-1, N).float()
        if res_row_mod:
            assert r.shape[0] == res_row_mod
            r = r[torch.arange(M) % res_row_mod]
        acc = acc + r
    return acc


# ------------------------------------------------------------------------------------------------- ops.*
def _dev_init(t: torch.Tensor) -> None:
    return None


def layernorm(x, gamma, beta, eps, out16=None, out32=None) -> None:
    _note("layernorm")
    assert x.dtype == torch.float32
    y = F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)
    if out32 is not None:
        out32.view(y.shape).copy_(y)
    if out16 is not None:
        out16.view(y.shape).copy_(y.to(out16.dtype))


def linear(a, w, *, bias=None, act=ACT_NONE, res_f32=None, res_16=None, out_f32=None, out_16=None, out_16_relu=None,
           block_n=0, res_row_mod=0, cta_pair=0, prof=None, group_rows=0, n_out=0) -> None:
    _note("linear")
    assert _is16(a) and a.dtype == w.dtype and a.is_contiguous() and w.is_contiguous()
    K = a.shape[-1]
    assert w.shape[1] == K
    if group_rows > 0:   # grouped weights: every group of rows multiplies its own block of the weight stack
        M = a.numel() // K
        groups = M // group_rows
        nw = w.shape[0] // groups
        acc = torch.bmm(a.reshape(groups, group_rows, K).float(), w.reshape(groups, nw, K).float().transpose(1, 2))
        acc = acc[..., :(n_out or nw)].reshape(M, -1)
    else:
        acc = a.reshape(-1, K).float() @ w.float().t()
    acc = _epilogue(acc, bias=bias, act=act, res_f32=res_f32, res_16=res_16, res_row_mod=res_row_mod)
    _store(acc, out_f32, out_16, out_16_relu)


def linear_qkv(a, w, bias, q, k, vt, heads, head_dim, tokens, block_n=0) -> None:
    _note("linear_qkv")
    assert _is16(a) and a.dtype == w.dtype == q.dtype
    K = a.shape[-1]
    acc = a.reshape(-1, K).float() @ w.float().t() + bias
    B = acc.shape[0] // tokens
    r = acc.reshape(B, tokens, 3, heads, head_dim).permute(2, 0, 3, 1, 4)   # [3,B,H,N,d]
    q[..., :head_dim] = r[0].to(q.dtype)
    k[..., :head_dim] = r[1].to(k.dtype)
    vt[:, :, :head_dim] = r[2].transpose(-1, -2).to(vt.dtype)


def _conv_weight(w_k, ksize, cin):
    cout = w_k.shape[0]
    kT, kH, kW = ksize
    return w_k.float().reshape(cout, kT, kH, kW, cin).permute(0, 4, 1, 2, 3).contiguous()


def conv3d(x, w, *, ksize, bias=None, act=ACT_NONE, res_16=None, res2_16=None, out_16=None, out_16_relu=None,
           out_f32=None, head_w2=None, head_b2=None, head_exp=False, block_n=0, cta_pair=0, prof=None, groups=1) -> None:
    _note("conv3d")
    assert _is16(x) and x.dtype == w.dtype and x.is_contiguous()
    B, T, H, W, Cin = x.shape
    kT, kH, kW = ksize
    assert w.shape[1] == kT * kH * kW * Cin
    if groups > 1:   # several heads' identical layers in one call: block g of the batch axis with its own weights / bias
        assert head_w2 is None and B % groups == 0 and w.shape[0] % groups == 0
        bg, co = B // groups, w.shape[0] // groups
        sl = lambda t, g, n: None if t is None else t.reshape(groups, -1)[g].reshape(n)
        for g in range(groups):
            part = lambda t: None if t is None else t.reshape(groups, bg, *t.shape[1:])[g]
            conv3d(x[g * bg:(g + 1) * bg].contiguous(), w[g * co:(g + 1) * co].contiguous(), ksize=ksize,
                   bias=None if bias is None else bias[g * co:(g + 1) * co], act=act, res_16=part(res_16), res2_16=part(res2_16),
                   out_16=part(out_16), out_16_relu=part(out_16_relu), out_f32=part(out_f32))
        return
    y = F.conv3d(x.float().permute(0, 4, 1, 2, 3), _conv_weight(w, ksize, Cin), None,
                 padding=(kT // 2, kH // 2, kW // 2)).permute(0, 2, 3, 4, 1)
    acc = y.reshape(-1, w.shape[0])
    if head_w2 is not None:
        hid = (acc + bias).clamp_min(0)
        o = hid @ head_w2.t() + head_b2                                       # [M, C2]
        if head_exp:
            o = torch.exp(o)
        out_f32.copy_(o.reshape(B, T, H, W, -1).permute(0, 4, 1, 2, 3))
        return
    acc = _epilogue(acc, bias=bias, act=act, res_16=res_16, res2_16=res2_16)
    _store(acc, out_f32, out_16, out_16_relu)


def _convT_weight(w, stride, cin):
    sT, sH, sW = stride
    cout = w.shape[0] // (sT * sH * sW)
    # rows (kt,kh,kw,co), cols Cin -> torch ConvTranspose3d layout [Cin,Cout,kt,kh,kw]
    return w.float().reshape(sT, sH, sW, cout, cin).permute(4, 3, 0, 1, 2).contiguous(), cout


def conv_transpose3d(x, w, bias, stride, out_16, block_n=0) -> None:
    _note("conv_transpose3d")
    assert _is16(x) and x.dtype == w.dtype == out_16.dtype
    B, T, H, W, Cin = x.shape
    wt, cout = _convT_weight(w, stride, Cin)
    y = F.conv_transpose3d(x.float().permute(0, 4, 1, 2, 3), wt, bias[:cout], stride=tuple(stride))
    out_16.copy_(y.permute(0, 2, 3, 4, 1).to(out_16.dtype))


def conv_transpose3d_hyper(x, w, bias, stride, hyper, out_f32, act=ACT_GELU, prof=None) -> None:
    _note("conv_transpose3d_hyper")
    G, T, H, W, Cin = x.shape
    wt, cout = _convT_weight(w, stride, Cin)
    up = _act(F.conv_transpose3d(x.float().permute(0, 4, 1, 2, 3), wt, bias[:cout], stride=tuple(stride)), act)
    out_f32.copy_(torch.einsum("gcthw,gkc->gkthw", up, hyper))


def attention(q, k, vt, out, head_dim, scale, prof=None) -> None:
    _note("attention")
    B, H, N, dpad = q.shape
    qf, kf, vf = q[..., :head_dim].float(), k[..., :head_dim].float(), vt[:, :, :head_dim].float().transpose(-1, -2)
    p = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1)
    o = p @ vf                                                                # [B,H,N,d]
    out.view(B, N, H, head_dim).copy_(o.permute(0, 2, 1, 3).to(out.dtype))


def patchify(rgb, out16, tubelet) -> None:
    _note("patchify")
    B, Cc, T, H, W = rgb.shape
    pt, ph, pw = tubelet
    x = rgb.reshape(B, Cc, T // pt, pt, H // ph, ph, W // pw, pw).permute(0, 2, 4, 6, 1, 3, 5, 7)
    out16.copy_(x.reshape(out16.shape).to(out16.dtype))


def cast16(x, y16, add=None) -> None:
    _note("cast16")
    assert x.dtype == torch.float32 and _is16(y16)
    y16.view(x.shape).copy_((x if add is None else x + add.view(x.shape)).to(y16.dtype))


def upsample3d(x, out_size, *, align_corners, y=None, y_relu=None) -> None:
    _note("upsample3d")
    r = F.interpolate(x.float().permute(0, 4, 1, 2, 3), size=tuple(out_size), mode="trilinear",
                      align_corners=bool(align_corners)).permute(0, 2, 3, 4, 1)
    if y is not None:
        y.copy_(r.to(y.dtype))
    if y_relu is not None:
        y_relu.copy_(r.clamp_min(0).to(y_relu.dtype))


def im2col3(x, out, stride) -> None:
    _note("im2col3")
    B, T, H, W, Cc = x.shape
    sT, sH, sW = stride
    xp = F.pad(x.float().permute(0, 4, 1, 2, 3), (1, 1, 1, 1, 1, 1))          # [B,C,T+2,H+2,W+2]
    cols = xp.unfold(2, 3, sT).unfold(3, 3, sH).unfold(4, 3, sW)             # [B,C,To,Ho,Wo,kt,kh,kw]
    cols = cols.permute(0, 2, 3, 4, 5, 6, 7, 1)                               # [B,To,Ho,Wo,kt,kh,kw,C]
    out.copy_(cols.reshape(out.shape).to(out.dtype))


def _sdpa(q, k, v, heads, scale):
    G, nq, C = q.shape
    d = C // heads
    qh = q.view(G, nq, heads, d).transpose(1, 2)
    kh = k.reshape(G, -1, heads, d).transpose(1, 2)
    vh = v.reshape(G, -1, heads, d).transpose(1, 2)
    a = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (a @ vh).transpose(1, 2).reshape(G, nq, C)


def token_attention(q, k16, v16, out, heads, shared_kv, scale) -> None:
    _note("token_attention")
    G, nq, C = q.shape
    Nk = k16.shape[0] if shared_kv else k16.shape[0] // G
    kf = k16.float().view(1 if shared_kv else G, Nk, C).expand(G, -1, -1)
    vf = v16.float().view(1 if shared_kv else G, Nk, C).expand(G, -1, -1)
    out.copy_(_sdpa(q, kf, vf, heads, scale))


def image_attention(q16, k, v, out16, G, heads, scale) -> None:
    _note("image_attention")
    Np = q16.shape[0] // G
    o = _sdpa(q16.float().view(G, Np, -1), k, v, heads, scale)
    out16.copy_(o.reshape(out16.shape).to(out16.dtype))


def layernorm16(x16, gamma, beta, eps, y16, gelu=False) -> None:
    _note("layernorm16")
    y = F.layer_norm(x16.float(), (x16.shape[-1],), gamma, beta, eps)
    if gelu:
        y = F.gelu(y)
    y16.copy_(y.to(y16.dtype))


def head_expand(q, out16, G, nt, heads, hd, scale) -> None:
    _note("head_expand")
    D = heads * hd
    qq = (q.reshape(G, 1, nt, heads, hd).float() * scale).expand(G, heads, nt, heads, hd)
    mask = torch.eye(heads, device=q.device).reshape(1, heads, 1, heads, 1)
    out16.view(G, heads, nt, heads, hd).copy_((qq * mask).to(out16.dtype))


def head_diag_gather(z, out, G, nt, heads, hd) -> None:
    _note("head_diag_gather")
    zz = z.reshape(G, heads, nt, heads, hd)
    idx = torch.arange(heads, device=z.device)
    out.view(G, nt, heads, hd).copy_(zz[:, idx, :, idx, :].permute(1, 2, 0, 3))   # [heads,G,nt,hd] -> [G,nt,heads,hd]


def row_softmax16(s, p16) -> None:
    _note("row_softmax16")
    assert s.dtype == torch.float32 and _is16(p16)
    p16.copy_(torch.softmax(s, dim=-1).to(p16.dtype))


def group_softmax_t16(s, p16, G, heads, nt) -> None:
    _note("group_softmax_t16")
    n = s.shape[-1]
    pr = torch.softmax(s.reshape(G, heads, nt, n), dim=2)             # over the tokens of a head
    p16.view(G, n, heads * nt).copy_(pr.reshape(G, heads * nt, n).transpose(1, 2).to(p16.dtype))


def token_weighted_sum(p16, x16, y16, G, J) -> None:
    _note("token_weighted_sum")
    n, C = p16.shape[-1], x16.shape[-1]
    y = torch.bmm(p16.reshape(G, J, n).float(), x16.reshape(G, n, C).float())
    y16.view(G, J, C).copy_(y.to(y16.dtype))


def track_readout(masks, image_hw):
    """sparse_heads.py:149-160 read-outs of the trilinearly upsampled mask logits (align_corners=False)."""
    _note("track_readout")
    G, nch, T, h, w = masks.shape
    H, W = image_hw
    logits = F.interpolate(masks, size=(T, H, W), mode="trilinear", align_corners=False)
    hm = torch.softmax(logits[:, 0].reshape(G, T, H * W), dim=-1)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5,
                            indexing="ij")
    traj = torch.stack([(hm * xs.reshape(-1)).sum(-1), (hm * ys.reshape(-1)).sum(-1)], dim=1)   # [G,2,T]
    vis = logits[:, 1].mean(dim=(-1, -2)).unsqueeze(1) if nch >= 2 else None
    depth = torch.exp(logits[:, 2].mean(dim=(-1, -2))).unsqueeze(1) if nch >= 3 else None
    return traj, vis, depth


# ------------------------------------------------------------------------------------ geometry wrappers
def _pose_call(camray_b6thw, k_norm, mode, output_size, thr, refits=3):
    """geometry_utils._pose_call: (extrinsics, pose = extrinsics^-1, centres, K estimate | None) via the oracle."""
    from oracle import l4p_oracle as O

    _note("pose_from_rays")
    rays = camray_b6thw.contiguous().float()
    centers = O.camera_centers(rays)
    if mode == 0:
        ext = O.rays_to_cameras(rays, k_norm.float())
        kest = None
    else:
        ext, kest = O.rays_to_cameras_fixed_intrinsics(rays, output_size=tuple(output_size), reproj_threshold=thr)[:2]
    pose = torch.linalg.inv(ext.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).contiguous()
    return ext, pose, centers, kest


def _affine_solve(self, pred, target, intrinsics=None, img_info=None, pred_conf=None, target_conf=None):
    from oracle import l4p_oracle as O

    _note("affine_align_solve")
    self.sol = O.lstsq_affine_solve(pred.float(), target.float(), inverse=bool(self.inverse))


def _affine_apply(self, pred):
    from oracle import l4p_oracle as O

    _note("affine_align_apply")
    return O.lstsq_affine_apply(self.sol, pred.float(), inverse=bool(self.inverse)).to(pred.dtype)


def _solve_sim3(pred, target, frame_step=3, rel_threshold=0.01, iters=8, min_points=10):
    """sim3.solve_sim3 stand-in: KabaschUmeyama3DAligner.solve (aligner.py:177-237) through the oracle restatement
    (all sampled-frame points instead of the random 10 % subset, seeded RANSAC)."""
    import numpy as np

    from oracle import l4p_oracle as O

    _note("sim3_align")
    d_s, d_t = pred["depth"].float(), target["depth"].float()
    bs, _, ov, H, W = d_s.shape
    assert bs == 1
    thr = float(torch.quantile(d_s.reshape(-1), 0.98)) * rel_threshold
    pts = []
    for side, d in ((pred, d_s), (target, d_t)):
        K = side["camray_intrinsics"].reshape(1, 4, 4, ov).float()[..., ::frame_step]
        P = side["camray"].reshape(1, 4, 4, ov).float()[..., ::frame_step]
        pts.append(O.generate_point_map(d[:, :, ::frame_step], K, P)[0].reshape(3, -1).T.double().numpy())
    T, _ = O.similarity_ransac(pts[0], pts[1], thr, min_samples=min_points)
    sim = O.similarity_from_T(T)
    dt = pred["depth"].dtype
    return {"T": torch.from_numpy(sim["T"])[None].to(dt), "s": torch.from_numpy(np.asarray(sim["s"])).reshape(1).to(dt),
            "R": torch.from_numpy(sim["R"])[None].to(dt), "t": torch.from_numpy(sim["t"])[None].to(dt)}


def install(monkeypatch) -> None:
    """Patch l4p_b200 so that the host mirror runs on CPU tensors through the functions above."""
    from l4p_b200 import ops
    from l4p_b200.models import aligner, sim3, videomae
    from l4p_b200.utils import geometry_utils

    CALLS.clear()
    for name in ("_dev_init", "layernorm", "linear", "linear_qkv", "conv3d", "conv_transpose3d", "conv_transpose3d_hyper",
                 "attention", "patchify", "cast16", "upsample3d", "im2col3", "token_attention", "image_attention",
                 "layernorm16", "track_readout", "head_expand", "head_diag_gather", "row_softmax16", "token_weighted_sum", "group_softmax_t16"):
        assert hasattr(ops, name), f"l4p_b200.ops.{name} no longer exists: update tests/emu.py"
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(geometry_utils, "_pose_call", _pose_call)
    monkeypatch.setattr(aligner.LstSqAffineAligner, "solve", _affine_solve)
    monkeypatch.setattr(aligner.LstSqAffineAligner, "apply", _affine_apply)
    monkeypatch.setattr(sim3, "solve_sim3", _solve_sim3)
    monkeypatch.setattr(videomae, "_require_device", lambda x: None)
