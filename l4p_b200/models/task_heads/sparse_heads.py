"""SAM-style point-tracking head (2D tracks + visibility + per-track depth) on the B200 kernels.

Drop-in mirror of l4p/models/task_heads/sparse_heads.py::VideoMAETrack2DSamHead (:19-667) and the SAM pieces it
owns (sam/prompt_encoder.py:19-232, sam/transformer.py:21-245, sam/mask_decoder.py:18-180, sam/common.py:16-31):
same constructor arguments, state-dict keys, `forward` / `forward_windowed` signatures and output keys.

Compute layout (B == 1, G = number of queries in the chunk, P = 2048 video tokens, C = 1408):
  * token side ([G,6,C], tiny): every Linear goes through the tcgen05 GEMM, LayerNorms through the LN kernel,
    the 6-token attentions through `token_attention`;
  * image side ([G*P, C], the 73.8 GF/query): k/v/q projections and out-projections are tcgen05 GEMMs with the
    "+ positional encoding" folded into a broadcast epilogue table (W·pe + b, indexed by row % P), the skinny
    attentions are the CUDA-core kernels in csrc/track.cu; while no per-query history exists (first window) the
    layer-0 k/v/q projections are computed ONCE for all queries instead of G times (the reference repeats the
    video tokens per query, mask_decoder.py:116-119);
  * mask decoder: ConvT(2,2,2) GEMM + pixel-shuffle, LayerNorm3d+GELU kernel, ConvT(1,2,2) GEMM whose epilogue
    applies GELU and the hyper-network dot, then the fused upsample + soft-argmax / mean read-out kernel.
"""
from __future__ import annotations

import math
from typing import List, Literal, Optional, Tuple

import torch
from torch import nn

from ... import lib as _l
from ... import ops
from ...utils.misc import apply_fn
from .. import params as P


# ------------------------------------------------------------------------------------------- containers
class _PE(nn.Module):
    def __init__(self, num_pos_feats: int, device=None):
        super().__init__()
        self.register_buffer("positional_encoding_gaussian_matrix", torch.randn((3, num_pos_feats), device=device))


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim, image_embedding_size, input_image_size, num_point_embeddings=2,
                 prompt_using_features=False, num_prompt_feature_embeddings=2, device=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = _PE(embed_dim // 2, device)
        self.num_point_embeddings = num_point_embeddings
        self.point_embeddings = nn.ModuleList([P.Embedding(1, embed_dim, device) for _ in range(num_point_embeddings)])
        self.prompt_using_features = prompt_using_features
        if prompt_using_features:
            self.prompt_feature_embeddings = nn.ModuleList(
                [P.Embedding(1, embed_dim, device) for _ in range(num_prompt_feature_embeddings)])
        self.not_a_point_embed = P.Embedding(1, embed_dim, device)
        self.no_mask_embed = P.Embedding(1, embed_dim, device)

    def _pe_encoding(self, coords01: torch.Tensor) -> torch.Tensor:
        """prompt_encoder.py:196-203 (coords in [0,1]^3, (t,x,y) order). Tiny: G x 2 points."""
        c = (2 * coords01 - 1) @ self.pe_layer.positional_encoding_gaussian_matrix.to(coords01)
        c = 2 * math.pi * c
        return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)

    def get_dense_pe(self) -> torch.Tensor:
        """prompt_encoder.py:205-219: cell-centre encoding of the (t,h,w) token grid -> [1,C,t,h,w]."""
        t, h, w = self.image_embedding_size
        dev = self.pe_layer.positional_encoding_gaussian_matrix.device
        grid = torch.ones((t, h, w), device=dev, dtype=torch.float32)
        te = (grid.cumsum(dim=0) - 0.5) / t
        ye = (grid.cumsum(dim=1) - 0.5) / h
        xe = (grid.cumsum(dim=2) - 0.5) / w
        return self._pe_encoding(torch.stack([te, xe, ye], dim=-1)).permute(3, 0, 1, 2).unsqueeze(0)

    def embed(self, coords_g13, labels_g1, feat_g1c, feat_labels_g) -> torch.Tensor:
        """forward(points=..., features=...) (prompt_encoder.py:78-180), mask-free: [G,3,C] =
        (point, 'not a point' pad, track feature)."""
        G = coords_g13.shape[0]
        T, H, W = self.input_image_size
        pts = torch.cat([coords_g13, torch.zeros(G, 1, 3, device=coords_g13.device, dtype=coords_g13.dtype)], dim=1)
        lab = torch.cat([labels_g1, -torch.ones(G, 1, device=labels_g1.device, dtype=labels_g1.dtype)], dim=1)
        # (t, x, y) extents, built on the device without a host->device copy (CUDA-graph capturable)
        scale = torch.stack([torch.full((), float(T), device=pts.device), torch.full((), float(W), device=pts.device),
                             torch.full((), float(H), device=pts.device)])
        pe = self._pe_encoding(pts.float() / scale)
        pe = torch.where((lab == -1)[..., None], self.not_a_point_embed.weight.expand_as(pe), pe)
        for i in range(self.num_point_embeddings):  # label 2 ("estimated") adds nothing when only 2 embeddings exist
            pe = pe + (lab == i)[..., None].to(pe.dtype) * self.point_embeddings[i].weight
        fe = torch.zeros_like(feat_g1c)
        fl = feat_labels_g.reshape(G, 1, 1)
        fe = fe + (fl == 0).to(fe.dtype) * (feat_g1c + self.prompt_feature_embeddings[0].weight)
        fe = fe + (fl == 1).to(fe.dtype) * (feat_g1c + self.prompt_feature_embeddings[1].weight)
        return torch.cat([pe, fe], dim=1)


class _Attention(nn.Module):
    def __init__(self, dim, heads, downsample_rate=1, device=None):
        super().__init__()
        self.embedding_dim, self.internal_dim, self.num_heads = dim, dim // downsample_rate, heads
        self.q_proj = P.Linear(dim, self.internal_dim, device=device)
        self.k_proj = P.Linear(dim, self.internal_dim, device=device)
        self.v_proj = P.Linear(dim, self.internal_dim, device=device)
        self.out_proj = P.Linear(self.internal_dim, dim, device=device)


class _MLPBlock(nn.Module):
    def __init__(self, dim, mlp_dim, device=None):
        super().__init__()
        self.lin1 = P.Linear(dim, mlp_dim, device=device)
        self.lin2 = P.Linear(mlp_dim, dim, device=device)


class TwoWayAttentionBlock(nn.Module):
    def __init__(self, dim, heads, mlp_dim, downsample, skip_first_layer_pe, device=None):
        super().__init__()
        self.self_attn = _Attention(dim, heads, device=device)
        self.norm1 = P.LayerNorm(dim, device=device)
        self.cross_attn_token_to_image = _Attention(dim, heads, downsample, device)
        self.norm2 = P.LayerNorm(dim, device=device)
        self.mlp = _MLPBlock(dim, mlp_dim, device)
        self.norm3 = P.LayerNorm(dim, device=device)
        self.norm4 = P.LayerNorm(dim, device=device)
        self.cross_attn_image_to_token = _Attention(dim, heads, downsample, device)
        self.skip_first_layer_pe = skip_first_layer_pe


class TwoWayTransformer(nn.Module):
    def __init__(self, depth, embedding_dim, num_heads, mlp_dim, attention_downsample_rate=2, device=None):
        super().__init__()
        self.depth, self.embedding_dim, self.num_heads, self.mlp_dim = depth, embedding_dim, num_heads, mlp_dim
        self.layers = nn.ModuleList([
            TwoWayAttentionBlock(embedding_dim, num_heads, mlp_dim, attention_downsample_rate, i == 0, device)
            for i in range(depth)])
        self.final_attn_token_to_image = _Attention(embedding_dim, num_heads, attention_downsample_rate, device)
        self.norm_final_attn = P.LayerNorm(embedding_dim, device=device)


class _MLP(nn.Module):
    def __init__(self, i, h, o, n, device=None):
        super().__init__()
        dims = [i] + [h] * (n - 1) + [o]
        self.layers = nn.ModuleList([P.Linear(a, b, device=device) for a, b in zip(dims[:-1], dims[1:])])


class _LayerNorm3d(nn.Module):
    def __init__(self, c, eps=1e-6, device=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.empty(c, device=device), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(c, device=device), requires_grad=False)


class MaskDecoder(nn.Module):
    def __init__(self, *, transformer_dim, transformer, num_mask_tokens=1, decoding_out_dim_factor=8, device=None):
        super().__init__()
        self.transformer_dim, self.transformer = transformer_dim, transformer
        self.iou_token = P.Embedding(1, transformer_dim, device)
        self.num_mask_tokens = num_mask_tokens
        self.mask_tokens = P.Embedding(num_mask_tokens, transformer_dim, device)
        d0 = min(2 * transformer_dim // decoding_out_dim_factor, transformer_dim)
        d1 = transformer_dim // decoding_out_dim_factor
        up = nn.Module()
        up.add_module("0", P.ConvTranspose3d(transformer_dim, d0, (2, 2, 2), (2, 2, 2), device))
        up.add_module("1", _LayerNorm3d(d0, device=device))
        up.add_module("3", P.ConvTranspose3d(d0, d1, (1, 2, 2), (1, 2, 2), device))
        self.output_upscaling = up
        self.output_hypernetworks_mlps = nn.ModuleList(
            [_MLP(transformer_dim, transformer_dim, d1, 3, device) for _ in range(num_mask_tokens)])


# ------------------------------------------------------------------------------------------- the head
class VideoMAETrack2DSamHead(nn.Module):
    compute_dtype = torch.float16
    # per-query video-token stream between the two-way layers: 16 bit (True, default: 5.2 instead of 10.4 GB of HBM traffic per
    # 128-query window, -0.63 ms, tracks equal to the fp32 stream within 6e-3 px) or fp32 (False); L4P_TRACK_STREAM16=0/1 overrides
    token_stream16 = __import__("os").environ.get("L4P_TRACK_STREAM16", "1") == "1"
    # token -> video-token attention with the K / V projections folded onto the 6-token side (csrc/track_t2i.cu);
    # L4P_TRACK_FOLD_T2I=0 restores the reference order (project all 2048 tokens of every query)
    fold_t2i = __import__("os").environ.get("L4P_TRACK_FOLD_T2I", "1") == "1"
    # the same for the video-token -> token attention (Q projection + 6-key attention + output projection of the token stream)
    fold_i2t = __import__("os").environ.get("L4P_TRACK_FOLD_I2T", "1") == "1"

    def __init__(self, task_name: str = "track_2d", prompt_embed_dim: int = 1408,
                 image_size: Tuple[int, int, int] = (16, 224, 224), patch_size: Tuple[int, int, int] = (2, 14, 14),
                 estimate_vis: bool = False, estimate_depth: bool = False, sam_head_depth: int = 2,
                 decoding_out_dim_factor: int = 8, num_prompt_points: int = 2, num_point_embeddings: int = 2,
                 modify_pointlabels_for_windowing: bool = False, prompt_using_features: bool = False,
                 attend_to_past: bool = False, depth_fn: str = "linear", vis_fn: str = "linear",
                 estimation_directions: List[Literal[1, -1]] = [1, -1], max_queries: int = 192, device=None):
        super().__init__()
        self.task_name, self.prompt_embed_dim = task_name, prompt_embed_dim
        self.image_size, self.patch_size = tuple(image_size), tuple(patch_size)
        self.estimate_vis, self.estimate_depth = estimate_vis, estimate_depth
        self.sam_head_depth, self.decoding_out_dim_factor = sam_head_depth, decoding_out_dim_factor
        self.num_prompt_points, self.num_point_embeddings = num_prompt_points, num_point_embeddings
        self.modify_pointlabels_for_windowing = modify_pointlabels_for_windowing
        self.prompt_using_features, self.attend_to_past = prompt_using_features, attend_to_past
        self.depth_fn, self.vis_fn = depth_fn, vis_fn
        self.estimation_directions, self.max_queries = estimation_directions, max_queries
        if not (prompt_using_features and attend_to_past and estimate_vis and estimate_depth):
            raise NotImplementedError("only the shipped configuration (estimate_vis/depth, prompt_using_features, "
                                      "attend_to_past: configs/model.yaml:53-66) is built")
        self.num_mask_tokens = 3
        self.token_ids = {"xy": 0, "vis": 1, "depth": 2, "prompt_feat": 3 + num_prompt_points}
        self.image_embedding_size = tuple(int(image_size[i] / patch_size[i]) for i in range(3))
        self.video_tokens_size = self.image_embedding_size[0] * self.image_embedding_size[1] * self.image_embedding_size[2]
        self.prompt_encoder = PromptEncoder(prompt_embed_dim, self.image_embedding_size, self.image_size,
                                            num_point_embeddings, True, device=device)
        self.mask_decoder = MaskDecoder(
            transformer=TwoWayTransformer(sam_head_depth, prompt_embed_dim, 8, 2048, device=device),
            transformer_dim=prompt_embed_dim, num_mask_tokens=self.num_mask_tokens,
            decoding_out_dim_factor=decoding_out_dim_factor, device=device)
        self.prompt_feature_linear_layer = P.Linear(prompt_embed_dim, prompt_embed_dim, device=device)
        self.processed_video_mask_token = P.Embedding(1, prompt_embed_dim, device)
        self.processed_video_features_proj = P.Linear(prompt_embed_dim, prompt_embed_dim, device=device)
        self.task_suffix = "_track_2d"
        # multi-GPU: shard the (independent) queries across the ranks of a process group (enable_query_sharding)
        self.shard_queries, self.query_shard_group = False, None
        self._packed = None
        self.register_load_state_dict_post_hook(lambda m, k: m.invalidate())

    def invalidate(self) -> None:
        self._packed = None

    def enable_query_sharding(self, enabled: bool = True, group=None) -> None:
        """Track only this rank's contiguous slice of the queries in `forward_windowed` and all-gather the tracks
        (every rank still needs the encoder features of all windows: combine with clip sharding, not window sharding)."""
        self.shard_queries, self.query_shard_group = bool(enabled), group

    # ------------------------------------------------------------------ weight packing
    def prepare(self, device, dt):
        if self._packed is not None and self._packed["device"] == device and self._packed["dtype"] == dt:
            return self._packed
        f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        w16 = lambda t: t.detach().to(device=device, dtype=dt).contiguous()
        C = self.prompt_embed_dim
        pk = {"device": device, "dtype": dt}
        pe = self.prompt_encoder.get_dense_pe().to(device)[0].reshape(C, -1).t().contiguous().float()  # [P, C]
        pk["pe"] = pe

        def lin(m):
            return dict(w=w16(m.weight), b=f(m.bias))

        def attn(a, image_k=False, image_q=False):
            d = dict(q=lin(a.q_proj), k=lin(a.k_proj), v=lin(a.v_proj), o=lin(a.out_proj), heads=a.num_heads,
                     hd=a.internal_dim // a.num_heads)
            # "+PE" folded: W(x + pe) + b = W x + (W pe + b); the table is added by the GEMM epilogue (row % P)
            if image_k:
                d["k_pe"] = (pe @ a.k_proj.weight.detach().to(device).float().t() + f(a.k_proj.bias)).contiguous()
                # operands of the folded token -> video-token attention (csrc/track_t2i.cu): W_k^T as a [C, inner] weight
                # (Q' = Q_blockdiag W_k) and the 16-bit copy of the positional / bias table (scores' position term)
                d["k_wT"] = w16(a.k_proj.weight.detach().t())
                d["k_pe16"] = d["k_pe"].to(dt).contiguous()
            if image_q:
                d["q_pe"] = (pe @ a.q_proj.weight.detach().to(device).float().t() + f(a.q_proj.bias)).contiguous()
                d["q_wT"] = w16(a.q_proj.weight.detach().t())     # folded video-token -> token attention (see i2t_folded)
                d["q_pe16"] = d["q_pe"].to(dt).contiguous()
            return d

        tr = self.mask_decoder.transformer
        layers = []
        for blk in tr.layers:
            layers.append(dict(
                sa=attn(blk.self_attn), t2i=attn(blk.cross_attn_token_to_image, image_k=True),
                i2t=attn(blk.cross_attn_image_to_token, image_q=True),
                lin1=lin(blk.mlp.lin1), lin2=lin(blk.mlp.lin2), skip_pe=blk.skip_first_layer_pe,
                n1=(f(blk.norm1.weight), f(blk.norm1.bias), blk.norm1.eps), n2=(f(blk.norm2.weight), f(blk.norm2.bias), blk.norm2.eps),
                n3=(f(blk.norm3.weight), f(blk.norm3.bias), blk.norm3.eps), n4=(f(blk.norm4.weight), f(blk.norm4.bias), blk.norm4.eps)))
        pk["layers"] = layers
        pk["final"] = attn(tr.final_attn_token_to_image, image_k=True)
        pk["nf"] = (f(tr.norm_final_attn.weight), f(tr.norm_final_attn.bias), tr.norm_final_attn.eps)
        up = self.mask_decoder.output_upscaling
        c0, ln, c3 = getattr(up, "0"), getattr(up, "1"), getattr(up, "3")

        def convT(m):
            st, sh, sw = m.stride
            return dict(w=w16(m.weight.detach().permute(2, 3, 4, 1, 0).reshape(st * sh * sw * m.out_channels, -1)),
                        b=f(m.bias).repeat(st * sh * sw).contiguous(), stride=m.stride, cout=m.out_channels)

        pk["up0"], pk["up3"] = convT(c0), convT(c3)
        pk["upln"] = (f(ln.weight), f(ln.bias), ln.eps)
        pk["hyper"] = [[lin(l) for l in mlp.layers] for mlp in self.mask_decoder.output_hypernetworks_mlps]
        pk["mask_tokens"] = f(self.mask_decoder.mask_tokens.weight)
        pk["pfl"] = lin(self.prompt_feature_linear_layer)
        pk["hist"] = lin(self.processed_video_features_proj)
        pk["hist_mask_token"] = f(self.processed_video_mask_token.weight[0])
        self._packed = pk
        return pk

    # ------------------------------------------------------------------ small helpers (token side)
    def _tok16(self, x32: torch.Tensor, add32: Optional[torch.Tensor] = None) -> torch.Tensor:
        """16-bit GEMM operand of fp32 rows, optionally of (x32 + add32) (the "+ positional embedding" of the attention inputs,
        one kernel instead of an add and a cast)."""
        rows = x32.numel() // x32.shape[-1]
        x16 = torch.empty(rows, x32.shape[-1], device=x32.device, dtype=self.compute_dtype)
        ops.cast16(x32.contiguous(), x16, add=None if add32 is None else add32.contiguous())
        return x16

    def _lin32(self, x32: Optional[torch.Tensor], l, act=_l.ACT_NONE, res32=None, out16=False, x16: Optional[torch.Tensor] = None):
        """Linear on fp32 rows through the tcgen05 GEMM: cast -> GEMM(+bias, act, +residual). A caller that feeds the same rows
        to several projections passes their 16-bit copy `x16` (from _tok16) instead of x32."""
        dt = self.compute_dtype
        if x16 is None:
            x16 = self._tok16(x32)
        rows, dev = x16.shape[0], x16.device
        N = l["w"].shape[0]
        if out16:
            y = torch.empty(rows, N, device=dev, dtype=dt)
            ops.linear(x16, l["w"], bias=l["b"], act=act, out_16=y)
        else:
            y = torch.empty(rows, N, device=dev, dtype=torch.float32)
            ops.linear(x16, l["w"], bias=l["b"], act=act, res_f32=res32, out_f32=y)
        return y

    def _ln32(self, x32: torch.Tensor, n) -> torch.Tensor:
        y = torch.empty_like(x32)
        ops.layernorm(x32.contiguous(), n[0], n[1], n[2], out32=y)
        return y

    def _token_attn(self, a, qk16, v16_in, G):
        """Attention among the prompt tokens themselves (6 keys), sam/transformer.py:157-161. qk16 / v16_in: 16-bit operand
        copies of the q = k input (queries + pe) and of the v input (queries)."""
        q = self._lin32(None, a["q"], x16=qk16).view(G, -1, a["q"]["w"].shape[0])
        k16 = self._lin32(None, a["k"], out16=True, x16=qk16)
        v16 = self._lin32(None, a["v"], out16=True, x16=v16_in)
        o = torch.empty_like(q)
        ops.token_attention(q, k16, v16, o, a["heads"], shared_kv=False, scale=1.0 / math.sqrt(a["hd"]))
        return o

    # ------------------------------------------------------------------ one window
    @torch.no_grad()
    def _decode(self, feat32, feat16, hist32, coords_g3, labels_g, pfeat_gc, plabels_g, need_history: bool):
        """forward_single_batch (sparse_heads.py:593-667) for G queries. feat32 [P,C] fp32, feat16 [P,C] 16-bit,
        hist32 [G,P,C] fp32 or None."""
        dev, dt = feat32.device, self.compute_dtype
        pk = self.prepare(dev, dt)
        G, C, Pn = coords_g3.shape[0], self.prompt_embed_dim, self.video_tokens_size
        sparse = self.prompt_encoder.embed(coords_g3.unsqueeze(1), labels_g.unsqueeze(1), pfeat_gc.unsqueeze(1), plabels_g)
        tokens = torch.cat([pk["mask_tokens"].unsqueeze(0).expand(G, -1, -1), sparse.float()], dim=1).contiguous()
        nt = tokens.shape[1]
        qpe = tokens                      # query_pe = the initial point embedding (transformer.py:92-101)
        queries = tokens

        shared = hist32 is None
        if shared:
            keys32, keys16 = feat32, feat16                      # [P,C]
        else:
            keys32 = (feat32.unsqueeze(0) + hist32).reshape(G * Pn, C).contiguous()
            keys16 = torch.empty(G * Pn, C, device=dev, dtype=dt)
            ops.cast16(keys32, keys16)

        def img_proj(x16, l, table):
            y = torch.empty(x16.shape[0], l["w"].shape[0], device=dev, dtype=dt)
            if table is None:
                ops.linear(x16, l["w"], bias=l["b"], out_16=y)
            else:
                ops.linear(x16, l["w"], res_f32=table, res_row_mod=Pn, out_16=y)
            return y

        def t2i_folded(a, queries, keys16):
            """Per-query video tokens (keys16 [G*P, C]): K / V are never formed. score = x . (W_k^T q) + q . (W_k pe + b_k),
            out = W_v (sum_n p_n x_n) + b_v (csrc/track_t2i.cu); 48 = heads x tokens score rows per query."""
            H, hd = a["heads"], a["hd"]
            D, J = H * hd, H * nt
            q = self._lin32(None, a["q"], x16=self._tok16(queries, qpe))               # [G*nt, D] fp32
            qb = torch.empty(G * J, D, device=dev, dtype=dt)
            ops.head_expand(q, qb, G, nt, H, hd, 1.0 / math.sqrt(hd))                  # block-diagonal, pre-scaled
            qp = torch.empty(G * J, C, device=dev, dtype=dt)
            ops.linear(qb, a["k_wT"], out_16=qp)                                       # Q' = W_k[h]^T q[g,t,h]
            sc = torch.empty(G * J, Pn, device=dev, dtype=torch.float32)
            ops.linear(qb, a["k_pe16"], out_f32=sc)                                    # q . (W_k pe_n + b_k)
            ops.linear(qp, keys16, res_f32=sc, out_f32=sc, group_rows=J)               # + Q'_g X_g^T (grouped weights)
            pr = torch.empty(G * J, Pn, device=dev, dtype=dt)
            ops.row_softmax16(sc, pr)
            y = torch.empty(G * J, C, device=dev, dtype=dt)
            ops.token_weighted_sum(pr, keys16, y, G, J)                                # sum_n p_n x_n
            z = torch.empty(G * J, D, device=dev, dtype=torch.float32)
            ops.linear(y, a["v"]["w"], bias=a["v"]["b"], out_f32=z)
            o = torch.empty(G * nt, D, device=dev, dtype=torch.float32)
            ops.head_diag_gather(z, o, G, nt, H, hd)
            return self._lin32(o, a["o"], res32=queries.view(-1, C).contiguous()).view(G, nt, C)

        def t2i(a, queries, keys16, shared_now):
            # kernel limits of the folded path (l4p_token_weighted_sum / l4p_row_softmax16); other shapes keep the reference order
            fits = a["heads"] * nt <= 48 and Pn % 64 == 0 and Pn <= 2048
            if self.fold_t2i and not shared_now and (fits or not keys16.is_cuda):
                return t2i_folded(a, queries, keys16)
            q = self._lin32(None, a["q"], x16=self._tok16(queries, qpe)).view(G, nt, -1)
            k16 = img_proj(keys16, a["k"], a["k_pe"])
            v16 = img_proj(keys16, a["v"], None)
            o = torch.empty_like(q)
            ops.token_attention(q, k16, v16, o, a["heads"], shared_kv=shared_now, scale=1.0 / math.sqrt(a["hd"]))
            return self._lin32(o.view(-1, o.shape[-1]), a["o"], res32=queries.view(-1, C).contiguous()).view(G, nt, C)

        n_layers = len(pk["layers"])
        for li, w in enumerate(pk["layers"]):
            # (1) token self attention
            if w["skip_pe"]:
                t16 = self._tok16(queries)
                o = self._token_attn(w["sa"], t16, t16, G)
                queries = self._lin32(o.view(-1, C), w["sa"]["o"]).view(G, nt, C)
            else:
                o = self._token_attn(w["sa"], self._tok16(queries, qpe), self._tok16(queries), G)
                queries = self._lin32(o.view(-1, C), w["sa"]["o"], res32=queries.view(-1, C).contiguous()).view(G, nt, C)
            queries = self._ln32(queries, w["n1"])
            # (2) tokens attend to the video tokens
            queries = self._ln32(t2i(w["t2i"], queries, keys16, shared), w["n2"])
            # (3) token MLP (ReLU, transformer.py:28,146)
            h16 = self._lin32(queries.view(-1, C), w["lin1"], act=_l.ACT_RELU, out16=True)
            m = torch.empty(G * nt, C, device=dev, dtype=torch.float32)
            ops.linear(h16, w["lin2"]["w"], bias=w["lin2"]["b"], res_f32=queries.view(-1, C).contiguous(), out_f32=m)
            queries = self._ln32(m.view(G, nt, C), w["n3"])
            # (4) video tokens attend to the prompt tokens -> per-query video tokens
            a = w["i2t"]
            fits = a["heads"] * nt <= 48 and a["heads"] * nt % 8 == 0 and Pn <= 2048
            if self.fold_i2t and self.token_stream16 and (fits or not keys16.is_cuda):
                # Folded form (csrc/track_t2i.cu): the Q projection of the 2048 video tokens of every query, the 6-key attention
                # and the output projection collapse into   S^T = (W_q^T k) X^T + k . (W_q pe + b_q)   (48 score rows per query),
                # a softmax over the 6 tokens of each head, and   new = P (W_o v) + b_o + residual   with a K = 48 GEMM per
                # query: the 704-wide Q and attention-output copies of the token stream are never formed.
                H, hd = a["heads"], a["hd"]
                D, J = H * hd, H * nt
                kk = self._lin32(None, a["k"], x16=self._tok16(queries, qpe))           # [G*nt, D] fp32
                vv = self._lin32(queries.view(-1, C), a["v"])
                kb = torch.empty(G * J, D, device=dev, dtype=dt)
                ops.head_expand(kk, kb, G, nt, H, hd, 1.0 / math.sqrt(hd))
                kp = torch.empty(G * J, C, device=dev, dtype=dt)
                ops.linear(kb, a["q_wT"], out_16=kp)                                    # K' = W_q[h]^T k[g,t,h]
                sc = torch.empty(G * J, Pn, device=dev, dtype=torch.float32)
                ops.linear(kb, a["q_pe16"], out_f32=sc)                                 # k . (W_q pe_n + b_q)
                if shared:
                    ops.linear(kp, keys16, res_f32=sc, out_f32=sc)                      # all queries still share the video tokens
                else:
                    ops.linear(kp, keys16, res_f32=sc, out_f32=sc, group_rows=J)
                p2 = torch.empty(G * Pn, J, device=dev, dtype=dt)
                ops.group_softmax_t16(sc, p2, G, H, nt)                                 # softmax over the tokens of a head, [G*P, J]
                vb = torch.empty(G * J, D, device=dev, dtype=dt)
                ops.head_expand(vv, vb, G, nt, H, hd, 1.0)
                vp = torch.empty(G * J, C, device=dev, dtype=dt)
                ops.linear(vb, a["o"]["w"], out_16=vp)                                  # V' = W_o[:, head h] v[g,t,h]
                # new = P (W_o v) + b_o + residual through the grouped K = 48 GEMM, then the 16-bit LayerNorm. (A single kernel
                # doing both with mma.sync - whole rows per warp and two passes over the columns, or column slices per warp
                # with the row moments merged through shared memory - was measured in round 2: correct, but 8.7 / 9.1 ms per
                # window against 8.1-8.3 ms for these two kernels; removed.)
                vpt = vp.view(G, J, C).transpose(1, 2).contiguous().view(G * C, J)      # per-query [C, J] weight block
                new16 = torch.empty(G * Pn, C, device=dev, dtype=dt)
                if keys32 is not None:
                    ops.linear(p2, vpt, bias=a["o"]["b"], res_f32=keys32, res_row_mod=Pn if shared else 0, out_16=new16, group_rows=Pn)
                else:
                    ops.linear(p2, vpt, bias=a["o"]["b"], res_16=keys16, out_16=new16, group_rows=Pn)
                nk16 = torch.empty(G * Pn, C, device=dev, dtype=dt)
                ops.layernorm16(new16, w["n4"][0], w["n4"][1], w["n4"][2], nk16)
                keys16 = nk16
                keys32 = None
                shared = False
                continue
            q16 = img_proj(keys16, a["q"], a["q_pe"])                                   # [rows, 704]
            k = self._lin32(None, a["k"], x16=self._tok16(queries, qpe)).view(G, nt, -1)
            v = self._lin32(queries.view(-1, C), a["v"]).view(G, nt, -1)
            if shared:  # first use of per-query state: the attention output differs per query
                q16 = q16.unsqueeze(0).expand(G, -1, -1).reshape(G * Pn, -1)
            ao = torch.empty(G * Pn, q16.shape[-1], device=dev, dtype=dt)
            ops.image_attention(q16.contiguous(), k.contiguous(), v.contiguous(), ao, G, a["heads"], 1.0 / math.sqrt(a["hd"]))
            if self.token_stream16:
                # 16-bit per-query video-token stream between the two-way layers: the out-projection adds the residual in
                # fp32 inside its epilogue (from the fp32 tokens in the first layer, from the previous layer's 16-bit LN
                # output afterwards) and stores 16 bit; the LayerNorm reads and writes 16 bit. Per 128-query call this moves
                # 5.2 GB instead of 10.4 GB through HBM (the fp32 stream is written, re-read by the LayerNorm, written again
                # as the next residual and re-read by the next out-projection).
                new16 = torch.empty(G * Pn, C, device=dev, dtype=dt)
                if keys32 is not None:
                    ops.linear(ao, a["o"]["w"], bias=a["o"]["b"], res_f32=keys32, res_row_mod=Pn if shared else 0, out_16=new16)
                else:
                    ops.linear(ao, a["o"]["w"], bias=a["o"]["b"], res_16=keys16, out_16=new16)
                keys16 = torch.empty(G * Pn, C, device=dev, dtype=dt)
                ops.layernorm16(new16, w["n4"][0], w["n4"][1], w["n4"][2], keys16)
                keys32 = None
                shared = False
                continue
            new32 = torch.empty(G * Pn, C, device=dev, dtype=torch.float32)
            ops.linear(ao, a["o"]["w"], bias=a["o"]["b"], res_f32=keys32, res_row_mod=Pn if shared else 0, out_f32=new32)
            keys16 = torch.empty(G * Pn, C, device=dev, dtype=dt)
            if li + 1 < n_layers:
                keys32 = new32   # the fp32 LN output is the next layer's residual (written in place)
                ops.layernorm(new32, w["n4"][0], w["n4"][1], w["n4"][2], out16=keys16, out32=keys32)
            else:
                keys32 = None    # after the last layer only the 16-bit operand copy is consumed (1.5 GB of HBM writes saved)
                ops.layernorm(new32, w["n4"][0], w["n4"][1], w["n4"][2], out16=keys16)
            shared = False
        # final token -> image attention (transformer.py:104-109)
        queries = self._ln32(t2i(pk["final"], queries, keys16, shared), pk["nf"])
        io = queries                                                               # [G,6,C]

        # hyper networks (mask_decoder.py:130-133)
        hyper = []
        for i in range(self.num_mask_tokens):
            x = io[:, i, :].contiguous()
            l0, l1, l2 = pk["hyper"][i]
            x16 = self._lin32(x, l0, act=_l.ACT_RELU, out16=True)
            y16 = torch.empty_like(x16)
            ops.linear(x16, l1["w"], bias=l1["b"], act=_l.ACT_RELU, out_16=y16)
            z = torch.empty(G, l2["w"].shape[0], device=dev, dtype=torch.float32)
            ops.linear(y16, l2["w"], bias=l2["b"], out_f32=z)
            hyper.append(z)
        hyper = torch.stack(hyper, dim=1).contiguous()                              # [G,3,176]

        # mask decoder upscaling (mask_decoder.py:58-66,135-139)
        et, eh, ew = self.image_embedding_size
        up0, up3 = pk["up0"], pk["up3"]
        s0 = up0["stride"]
        u1 = torch.empty(G, et * s0[0], eh * s0[1], ew * s0[2], up0["cout"], device=dev, dtype=dt)
        ops.conv_transpose3d(keys16.view(G, et, eh, ew, C), up0["w"], up0["b"], s0, u1)
        u1n = torch.empty_like(u1)
        ops.layernorm16(u1, pk["upln"][0], pk["upln"][1], pk["upln"][2], u1n, gelu=True)
        s3 = up3["stride"]
        masks = torch.empty(G, self.num_mask_tokens, u1.shape[1] * s3[0], u1.shape[2] * s3[1], u1.shape[3] * s3[2],
                            device=dev, dtype=torch.float32)
        ops.conv_transpose3d_hyper(u1n, up3["w"], up3["b"], s3, hyper, masks, act=_l.ACT_GELU)
        traj, vis, depth = ops.track_readout(masks, (self.image_size[1], self.image_size[2]))

        pfeat = self._lin32(io[:, self.token_ids["prompt_feat"], :].contiguous(), pk["pfl"])   # [G,C]
        hist = None
        if need_history:
            hist = torch.empty(G * Pn, C, device=dev, dtype=torch.float32)
            ops.linear(keys16, pk["hist"]["w"], bias=pk["hist"]["b"], out_f32=hist)
            hist = hist.view(G, Pn, C)
        return traj, vis, depth, pfeat, hist

    def forward(self, enc_features_bpc_list, track_2d_pointquerries_bn3: torch.Tensor,
                track_2d_pointlabels_bn: torch.Tensor, track_2d_promptfeatures_bnc: Optional[torch.Tensor] = None,
                track_2d_promptfeaturelabels_bn: Optional[torch.Tensor] = None, _need_history: bool = True, **kwargs):
        """Single-window tracking (sparse_heads.py:497-591). enc_features_bpc_list[-1]: [B,P,C] or [B,N,P,C]."""
        enc = enc_features_bpc_list[-1]
        B = enc.shape[0]
        outs = {k: [] for k in ("traj", "vis", "depth", "pfeat", "hist")}
        cached16 = getattr(enc_features_bpc_list, "taps16", None)
        for b in range(B):
            q = track_2d_pointquerries_bn3[b].float()
            G = q.shape[0]
            lab = track_2d_pointlabels_bn[b].float()
            C = self.prompt_embed_dim
            pf = (track_2d_promptfeatures_bnc[b].float() if track_2d_promptfeatures_bnc is not None
                  else torch.zeros(G, C, device=q.device))
            pl = (track_2d_promptfeaturelabels_bn[b].float() if track_2d_promptfeaturelabels_bn is not None
                  else torch.zeros(G, device=q.device))
            if enc.dim() == 3:
                feat32 = enc[b].contiguous().float()
                hist32 = None
                L = len(enc_features_bpc_list) - 1
                if cached16 is not None and L in cached16 and cached16[L].dtype == self.compute_dtype:
                    Pn = feat32.shape[0]
                    feat16 = cached16[L][b * Pn:(b + 1) * Pn]
                else:
                    feat16 = torch.empty(feat32.shape, device=feat32.device, dtype=self.compute_dtype)
                    ops.cast16(feat32, feat16)
                r = self._decode(feat32, feat16, None, q, lab, pf, pl, _need_history)
            else:  # [B,N,P,C]: encoder tokens already combined with the per-query history by the windowed driver
                keys = enc[b].contiguous().float()
                r = self._decode_with_keys(keys, q, lab, pf, pl, _need_history)
            for k, v in zip(outs, r):
                outs[k].append(v)
        out = {f"{self.task_name}_prompt_features_bnc": torch.stack(outs["pfeat"], 0)}
        if outs["hist"][0] is not None:
            out[f"{self.task_name}_enc_features_with_track_history_bnpc"] = torch.stack(outs["hist"], 0)
        out[f"{self.task_name}_traj_est_bn2t"] = torch.stack(outs["traj"], 0)
        out[f"{self.task_name}_vis_est_bn1t"] = apply_fn(torch.stack(outs["vis"], 0), self.vis_fn)
        depth = torch.stack(outs["depth"], 0)  # the kernel applied exp (depth_fn == 'exp', configs/model.yaml:65)
        if self.depth_fn != "exp":
            depth = apply_fn(torch.log(depth), self.depth_fn)
        out[f"{self.task_name}_depth_est_bn1t"] = depth
        return out

    def _decode_with_keys(self, keys_gpc, q, lab, pf, pl, need_history):
        """Per-query keys given explicitly ([G,P,C] = encoder tokens + history)."""
        G, Pn, C = keys_gpc.shape
        zero = torch.zeros(Pn, C, device=keys_gpc.device, dtype=torch.float32)
        z16 = torch.zeros(Pn, C, device=keys_gpc.device, dtype=self.compute_dtype)
        return self._decode(zero, z16, keys_gpc, q, lab, pf, pl, need_history)

    # ------------------------------------------------------------------ sliding windows
    def forward_windowed(self, enc_features_bpc_2dlist, track_2d_pointquerries_bn3: torch.Tensor,
                         track_2d_pointlabels_bn: torch.Tensor, time_strides: Optional[torch.Tensor] = None, **kwargs):
        """Query chunking by max_queries (sparse_heads.py:162-211)."""
        kwargs.pop("_batched_windows", None)
        N = track_2d_pointquerries_bn3.shape[1]
        if self.shard_queries:
            # multi-GPU: queries are independent -> this rank tracks its contiguous slice (chunked by
            # max_queries as usual), ONE all-gather returns the tracks of all queries on every rank
            import torch.distributed as dist
            from ...parallel import contiguous_partition, gather_query_outputs
            assert dist.is_initialized(), "query sharding needs an initialised torch.distributed process group"
            grp = self.query_shard_group
            start, count = contiguous_partition(N, dist.get_world_size(grp))[dist.get_rank(grp)]
            keys = [f"{self.task_name}_traj_est_bn2t", f"{self.task_name}_vis_est_bn1t", f"{self.task_name}_depth_est_bn1t"]
            local = [None] * len(keys)
            if count > 0:
                self.shard_queries = False
                try:
                    mine = self.forward_windowed(enc_features_bpc_2dlist, track_2d_pointquerries_bn3[:, start:start + count],
                                                 track_2d_pointlabels_bn[:, start:start + count], time_strides, **kwargs)
                finally:
                    self.shard_queries = True
                local = [mine[k] for k in keys]
            return dict(zip(keys, gather_query_outputs(local, N, grp)))
        if N < self.max_queries:
            return self.forward_windowed_core(enc_features_bpc_2dlist, track_2d_pointquerries_bn3,
                                              track_2d_pointlabels_bn, time_strides, **kwargs)
        out_list = []
        for i in range(int(math.ceil(N / self.max_queries))):
            sl = slice(i * self.max_queries, (i + 1) * self.max_queries)
            out_list.append(self.forward_windowed_core(enc_features_bpc_2dlist, track_2d_pointquerries_bn3[:, sl],
                                                       track_2d_pointlabels_bn[:, sl], time_strides, **kwargs))
        return {k: torch.cat([o[k] for o in out_list], dim=1) for k in out_list[0].keys()}

    def forward_windowed_core(self, enc_features_bpc_2dlist, track_2d_pointquerries_bn3, track_2d_pointlabels_bn,
                              time_strides=None, **kwargs):
        """Sliding-window memory tracker (sparse_heads.py:213-495), forward direction only (:242-245).
        Same state machine, with the per-query Python loops replaced by device-side gathers."""
        if time_strides is None:
            return self.forward(enc_features_bpc_2dlist[0], track_2d_pointquerries_bn3, track_2d_pointlabels_bn)
        dtype, device = track_2d_pointquerries_bn3.dtype, track_2d_pointquerries_bn3.device
        Tw = self.image_size[0]
        B, N = track_2d_pointquerries_bn3.shape[:2]
        T = int(time_strides[-1] + Tw)
        traj = torch.zeros(B, N, 2, T, dtype=dtype, device=device)
        vis = -torch.ones(B, N, 1, T, dtype=dtype, device=device) * 10.0
        depth = torch.zeros(B, N, 1, T, dtype=dtype, device=device)
        assert B == 1, "Currently only supports batch size of 1"
        assert len(self.estimation_directions) == 1 and self.estimation_directions[0] == 1, (
            "Currently only positive direction estimation is supported for sliding window tracking."
            "Run twice, with and without video flipping, and then combine outputs.")
        C, Pn = self.prompt_embed_dim, self.video_tokens_size
        et, eh, ew = self.image_embedding_size
        pfeat = torch.zeros(B, N, C, dtype=dtype, device=device)
        pfeat_lab = torch.zeros(B, N, dtype=dtype, device=device)
        hist = None  # None == the learned mask token everywhere (first window)
        cur_q = track_2d_pointquerries_bn3.clone()
        cur_lab = track_2d_pointlabels_bn.clone()
        nW = time_strides.shape[0]
        ar = torch.arange(Tw, device=device)
        for win_id in range(nW):
            s = int(time_strides[win_id])
            nxt = int(time_strides[win_id + 1]) if win_id < nW - 1 else int(time_strides[win_id - 1])
            q_off = cur_q.clone()
            valid_bn1t = ((ar.view(1, 1, Tw) + s + 0.5 - q_off[:, :, 0:1]) >= 0)[:, :, None, :]   # [B,N,1,Tw]
            valid_bn = valid_bn1t.sum(dim=-1)[..., 0] > 0
            q_off[:, :, 0] -= s
            cur_lab = torch.where(valid_bn, torch.ones_like(cur_lab), torch.zeros_like(cur_lab))
            if self.modify_pointlabels_for_windowing:
                same = (cur_q == track_2d_pointquerries_bn3).sum(dim=-1) > 0
                cur_lab = torch.where(same, torch.ones_like(cur_lab), cur_lab)
                cur_lab = torch.where(torch.logical_and(valid_bn, ~same), torch.full_like(cur_lab, 2), cur_lab)
            feats = enc_features_bpc_2dlist[win_id]
            last = win_id == nW - 1
            if hist is None and win_id == 0:
                # history == mask token for every video token of every query: fold it in as per-query keys only if
                # it is not all equal... it is identical for all queries, so add it to the shared tokens instead.
                enc_last = feats[-1] + self.processed_video_mask_token.weight[0].to(feats[-1])
                out = self.forward(_WithLast(feats, enc_last), q_off, cur_lab, pfeat, pfeat_lab, _need_history=not last)
            else:
                enc_last = feats[-1].unsqueeze(1) + hist                                       # [B,N,P,C]
                out = self.forward([enc_last], q_off, cur_lab, pfeat, pfeat_lab, _need_history=not last)
            sl = slice(s, s + Tw)
            vis[..., sl] = torch.where(valid_bn1t, out[f"{self.task_name}_vis_est_bn1t"].to(dtype), vis[..., sl])
            traj[..., sl] = torch.where(valid_bn1t, out[f"{self.task_name}_traj_est_bn2t"].to(dtype), traj[..., sl])
            depth[..., sl] = torch.where(valid_bn1t, out[f"{self.task_name}_depth_est_bn1t"].to(dtype), depth[..., sl])
            if last:
                continue
            pfeat = torch.where(valid_bn[..., None], out[f"{self.task_name}_prompt_features_bnc"].to(dtype), pfeat)
            pfeat_lab = torch.where(valid_bn, torch.ones_like(pfeat_lab), pfeat_lab)
            # memory: keep the second half (in time) of the decoded tokens, pad the rest with the mask token
            h = out[f"{self.task_name}_enc_features_with_track_history_bnpc"].to(dtype).view(B, N, et, eh * ew, C)
            mask_tok = self.processed_video_mask_token.weight[0].to(h).view(1, 1, 1, 1, C).expand(B, N, et // 2, eh * ew, C)
            hist = torch.cat([h[:, :, et // 2:], mask_tok], dim=2).reshape(B, N, Pn, C)
            # re-query every track at its most visible frame inside the overlap with the next window
            ov0, ov1 = nxt, s + Tw
            vis_ov = vis[..., ov0:ov1]
            best = torch.argmax(vis_ov, dim=-1)                                               # [B,N,1]
            xy = torch.gather(traj[..., ov0:ov1], 3, best[..., None].expand(B, N, 2, 1))[..., 0]  # [B,N,2]
            new_q = torch.cat([(best.to(dtype) + nxt + 0.5), xy], dim=-1)                       # (t, x, y)
            later = new_q[:, :, 0] > cur_q[:, :, 0]
            cur_q = torch.where(later[..., None], new_q, cur_q)
        return {f"{self.task_name}_traj_est_bn2t": traj, f"{self.task_name}_vis_est_bn1t": vis,
                f"{self.task_name}_depth_est_bn1t": depth}


class _WithLast(list):
    """A feature list whose last entry is replaced (keeps the encoder's 16-bit cache out of play)."""

    def __init__(self, feats, last):
        super().__init__(list(feats[:-1]) + [last])
