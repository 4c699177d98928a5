"""Folded token -> video-token attention of the track head (csrc/track_t2i.cu, grouped-weight l4p_gemm): every new kernel
against its plain-torch definition (tests/emu.py, the CPU stand-in the host tests use), and the folded head against the
reference order of operations (project all 2048 video tokens of every query) on the device.
Reference: l4p/models/task_heads/sam/transformer.py:223-245 (Attention.forward), :157-187 (TwoWayAttentionBlock)."""
import pytest
import torch

from tests import emu
from tests.util import grid_queries

pytestmark = pytest.mark.gpu
DT = [torch.float16, torch.bfloat16]


def _ops():
    from l4p_b200 import ops
    return ops


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


def _close(got, ref, tol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    assert err <= tol * max(ref.abs().max().item(), 1e-6), f"max abs err {err} vs max |ref| {ref.abs().max().item()}"


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("G,J,n,K", [(3, 48, 2048, 1408), (5, 48, 256, 704), (2, 16, 128, 192), (130, 48, 64, 64),
                                     (6, 48, 64, 704), (4, 128, 96, 1408)])   # narrow tiles: 128-wide K stages with grouped weights
def test_grouped_linear(dtype, G, J, n, K):
    """l4p_gemm with grouped weights: rows [g*J, (g+1)*J) of A against W rows [g*n, (g+1)*n); in-place fp32 residual; tiles
    J rows apart (J < 128: the rows a tile reads beyond its group are never stored)."""
    ops = _ops()
    a = _rand((G * J, K), dtype, 1)
    w = _rand((G * n, K), dtype, 2, K ** -0.5)
    res = _rand((G * J, n), torch.float32, 3)
    out = res.clone()
    ops.linear(a, w, res_f32=out, out_f32=out, group_rows=J)
    torch.cuda.synchronize()
    ref = torch.bmm(a.view(G, J, K).float(), w.view(G, n, K).float().transpose(1, 2)).reshape(G * J, n) + res
    _close(out, ref, 2e-5)
    o16 = torch.empty(G * J, n, device="cuda", dtype=dtype)
    ops.linear(a, w, out_16=o16, group_rows=J)
    torch.cuda.synchronize()
    _close(o16, ref - res, 2 ** -7 if dtype == torch.bfloat16 else 2 ** -10)


@pytest.mark.parametrize("dtype", DT)
def test_grouped_linear_long_groups(dtype):
    """Groups longer than a tile (2048 rows each), per-group weights with K = 48 (< one 64-wide k-block: TMA zero fill)."""
    ops = _ops()
    G, R, N, K = 3, 2048, 1408, 48
    a = _rand((G * R, K), dtype, 4)
    w = _rand((G * N, K), dtype, 5, K ** -0.5)
    b = _rand((N,), torch.float32, 6)
    out = torch.empty(G * R, N, device="cuda", dtype=dtype)
    ops.linear(a, w, bias=b, out_16=out, group_rows=R)
    torch.cuda.synchronize()
    ref = torch.bmm(a.view(G, R, K).float(), w.view(G, N, K).float().transpose(1, 2)).reshape(G * R, N) + b
    _close(out, ref, 2 ** -7 if dtype == torch.bfloat16 else 2 ** -10)


@pytest.mark.parametrize("dtype", DT)
def test_head_expand_and_gather(dtype):
    ops = _ops()
    G, nt, H, hd = 7, 6, 8, 88
    q = _rand((G * nt, H * hd), torch.float32, 7)
    out = torch.full((G * H * nt, H * hd), 3.0, device="cuda", dtype=dtype)
    ops.head_expand(q, out, G, nt, H, hd, 0.25)
    ref = torch.empty_like(out)
    emu.head_expand(q, ref, G, nt, H, hd, 0.25)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    z = _rand((G * H * nt, H * hd), torch.float32, 8)
    o = torch.empty(G * nt, H * hd, device="cuda")
    ops.head_diag_gather(z, o, G, nt, H, hd)
    r = torch.empty_like(o)
    emu.head_diag_gather(z, r, G, nt, H, hd)
    torch.cuda.synchronize()
    assert torch.equal(o, r)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,n", [(48 * 3, 2048), (13, 64), (9, 1000)])
def test_row_softmax16(dtype, rows, n):
    ops = _ops()
    s = _rand((rows, n), torch.float32, 9, 3.0)
    s[0, 5] = 40.0   # one peaked row
    p = torch.empty(rows, n, device="cuda", dtype=dtype)
    ops.row_softmax16(s, p)
    torch.cuda.synchronize()
    ref = torch.softmax(s, dim=-1)
    assert (p.float() - ref).abs().max().item() <= (2 ** -8 if dtype == torch.bfloat16 else 2 ** -11) * 1.01
    assert (p.float().sum(-1) - 1).abs().max().item() < 2e-2


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("G,J,n,C", [(3, 48, 2048, 1408), (2, 12, 128, 192), (4, 48, 64, 704), (1, 5, 192, 80)])
def test_token_weighted_sum(dtype, G, J, n, C):
    """Y[g] = P[g] X[g] (mma.sync streaming kernel) vs fp32 bmm of the same 16-bit operands; J < 48 (rows of the next query in
    the tile), channel counts that end inside a 128-wide slice."""
    ops = _ops()
    p = torch.softmax(_rand((G * J, n), torch.float32, 10, 2.0), dim=-1).to(dtype)
    x = _rand((G * n, C), dtype, 11)
    y = torch.full((G * J + 3, C), 5.0, device="cuda", dtype=dtype)
    ops.token_weighted_sum(p, x, y[:G * J], G, J)
    torch.cuda.synchronize()
    ref = torch.bmm(p.view(G, J, n).float(), x.view(G, n, C).float()).reshape(G * J, C)
    _close(y[:G * J], ref, 2 ** -7 if dtype == torch.bfloat16 else 2 ** -10)
    assert bool((y[G * J:] == 5.0).all())


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("G,H,nt,n", [(3, 8, 6, 2048), (2, 2, 5, 100), (1, 4, 1, 64)])
def test_group_softmax_t16(dtype, G, H, nt, n):
    ops = _ops()
    s = _rand((G * H * nt, n), torch.float32, 12, 3.0)
    p = torch.empty(G * n, H * nt, device="cuda", dtype=dtype)
    ops.group_softmax_t16(s, p, G, H, nt)
    ref = torch.empty_like(p)
    emu.group_softmax_t16(s, ref, G, H, nt)
    torch.cuda.synchronize()
    assert (p.float() - ref.float()).abs().max().item() <= (2 ** -8 if dtype == torch.bfloat16 else 2 ** -11) * 1.01


def test_folded_head_equals_reference_order():
    """The whole track head at the bench's size class (full token size, 16 queries, per-query tokens after the first layer)
    with the folded token -> video-token attention against the same head projecting K / V of all video tokens."""
    from l4p_b200 import weights
    from l4p_b200.models.task_heads.sparse_heads import VideoMAETrack2DSamHead

    h = VideoMAETrack2DSamHead(task_name="track_2d", estimate_vis=True, estimate_depth=True, sam_head_depth=2, num_point_embeddings=2,
                               prompt_using_features=True, attend_to_past=True, modify_pointlabels_for_windowing=True,
                               estimation_directions=[1], depth_fn="exp", vis_fn="linear").cuda()
    weights.fill_module_fast_(h, seed=3)
    g = torch.Generator(device="cuda").manual_seed(5)
    feat = torch.randn(1, 2048, 1408, device="cuda", generator=g)
    q = grid_queries(4).cuda()
    lab = torch.ones(1, q.shape[1], device="cuda")
    feats = [None] * 40 + [feat]
    outs = {}
    for fold in ((False, False), (True, False), (False, True), (True, True)):
        h.fold_t2i, h.fold_i2t = fold
        with torch.no_grad():
            outs[fold] = {k: v.float().clone() for k, v in h.forward_windowed([feats], q, lab, time_strides=torch.tensor([0])).items()}
    torch.cuda.synchronize()
    a = outs[(False, False)]
    for fold in ((True, False), (False, True), (True, True)):
        b = outs[fold]
        assert (a["track_2d_traj_est_bn2t"] - b["track_2d_traj_est_bn2t"]).abs().max().item() < 0.05, fold      # pixels
        assert (a["track_2d_vis_est_bn1t"] - b["track_2d_vis_est_bn1t"]).abs().max().item() < 5e-3, fold
        d = (a["track_2d_depth_est_bn1t"] - b["track_2d_depth_est_bn1t"]).abs() / a["track_2d_depth_est_bn1t"].abs()
        assert d.max().item() < 5e-3, fold
