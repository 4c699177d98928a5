/*
 * l4p_b200 — C ABI of the Blackwell-native (sm_100a) L4P inference hot path.
 *
 * The reference (NVlabs/L4P, pure Python/PyTorch) has no FFI layer: every "kernel" is an ATen
 * library call made from an nn.Module.forward. This header is therefore the boundary
 * a maintainer binds with ctypes/cffi from the module that used to make the ATen call; each entry
 * point cites the reference call site it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - "16-bit" tensors are fp16 or bf16 as selected by `bf16` (0 = fp16, 1 = bf16); accumulation,
 *     residual stream, LayerNorm/softmax statistics are always fp32;
 *   - sizes are int64_t / int, streams are passed as void* (cudaStream_t);
 *   - functions return 0 on success, <0 on error (L4P_ERR_*), never exit/throw; the message for the
 *     last error on the calling thread is l4p_last_error();
 *   - kernels never allocate: callers own all buffers (workspace sizes via *_workspace_bytes()).
 */
#ifndef L4P_B200_H_
#define L4P_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L4P_OK 0
#define L4P_ERR_ARG (-1)
#define L4P_ERR_SHAPE (-2)
#define L4P_ERR_ARCH (-3)
#define L4P_ERR_CUDA (-4)
#define L4P_ERR_DRIVER (-5)

/* ---- library ------------------------------------------------------------------------------- */
int l4p_version(void);
const char* l4p_last_error(void);
/* Select `device`, verify it is compute capability 10.x when arch_check != 0.
 * Replaces the implicit device placement of Fabric.setup (l4p/models/utils.py:57-58). */
int l4p_init(int device, int arch_check);

/* ---- K2: LayerNorm ------------------------------------------------------------------------- */
/* y16[r, :] = (x[r, :] - mean) * rstd * gamma + beta, biased variance, fp32 statistics.
 * Replaces Block.norm1/norm2 (l4p/models/VideoMAEv2/models/modeling_finetune.py:247-248), the final
 * encoder norm (l4p/models/l4p_videomae.py:115) and the SAM LayerNorms (task_heads/sam/transformer.py:139-149).
 * x fp32 [rows, cols]; y16 16-bit and/or y32 fp32 outputs (either may be NULL). */
int l4p_layernorm(const float* x, const float* gamma, const float* beta, void* y16, float* y32,
                  int64_t rows, int cols, float eps, int bf16, void* stream);

/* ---- K3/K5/K6/K7/K8: tcgen05 GEMM / implicit-GEMM convolution ---------------------------------- */
enum { L4P_ACT_NONE = 0, L4P_ACT_GELU = 1, L4P_ACT_RELU = 2, L4P_ACT_EXP = 3 };
enum {
  L4P_STORE_ROWMAJOR = 0, /* out[row, col]                                                        */
  L4P_STORE_QKV = 1,      /* scatter to Q[B,H,N,dpad], K[B,H,N,dpad], Vt[B,H,dpad,N]                */
  L4P_STORE_CONVT = 2,    /* ConvTranspose3d(k == s) pixel-shuffle scatter to channels-last         */
  L4P_STORE_HEAD1X1 = 3,  /* ReLU -> 1x1x1 conv to <= 8 channels (+exp) -> fp32 NCTHW              */
  L4P_STORE_HYPER = 4     /* ConvT(k==s) tap -> act -> dot with per-group hyper vectors -> fp32 masks */
};
enum { L4P_A_MATRIX = 0, L4P_A_CONV3D = 1 };

typedef struct l4p_gemm_desc {
  /* D[M,N] = A[M,K] * W[N,K]^T ; A and W 16-bit, K contiguous (nn.Linear weight layout). */
  const void* a;       /* A_MATRIX: [M, lda] ; A_CONV3D: channels-last activations [B,T,H,W,Cin]   */
  const void* w;       /* [N, ldw]; for A_CONV3D the K axis is (tap_t, tap_h, tap_w, Cin)           */
  int64_t M, N, K;     /* A_CONV3D: M = B*T*H*W output voxels, K = taps*Cin                        */
  int64_t lda, ldw;    /* row strides in elements (multiples of 8)                                 */
  int bf16;            /* operand type                                                             */
  int a_mode;          /* L4P_A_*                                                                  */
  /* A_CONV3D geometry (stride 1, zero padding k/2): */
  int cB, cT, cH, cW, cCin;
  int kT, kH, kW;      /* filter extent (odd)                                                      */
  int bT, bH, bW;      /* voxel box of one 128-row tile, bT*bH*bW == 128                           */
  /* epilogue */
  const float* bias;   /* [N] fp32 or NULL                                                         */
  int act;             /* L4P_ACT_* applied after bias                                             */
  const float* res_f32;   /* optional residual, same indexing as out_f32                           */
  const void* res_16;     /* optional 16-bit residual, same indexing as out_16 (row-major)         */
  const void* res2_16;    /* optional second 16-bit residual                                       */
  int64_t ld_res;
  int res_row_mod;        /* >0: res_f32 row index = row % res_row_mod (broadcast table, e.g. pos-embed) */
  int store_mode;      /* L4P_STORE_*                                                              */
  float* out_f32;      /* ROWMAJOR: optional fp32 output [M, ld_out]                               */
  void* out_16;        /* ROWMAJOR/CONVT: 16-bit output                                            */
  void* out_16_relu;   /* ROWMAJOR: optional second 16-bit output = relu(out)                      */
  int64_t ld_out;
  /* STORE_QKV: N = 3*heads*head_dim, rows = (b, token) */
  void* q; void* k; void* vt;
  int heads, head_dim, head_dim_pad, tokens;
  /* STORE_CONVT: rows are input voxels (b,t,h,w) of a [cB,cT,cH,cW] grid, N = sT*sH*sW*Cout ordered
   * (kt,kh,kw,co); out_16 is channels-last [B, T*sT, H*sH, W*sW, Cout] */
  int sT, sH, sW, ctCout;
  /* STORE_HEAD1X1: out_f32[b, c, t, h, w] = f(sum_n relu(acc+bias)[n] * w2[c, n] + b2[c]) */
  const float* w2; const float* b2; int c2; int exp_out;
  /* STORE_HYPER (mask decoder, sam/mask_decoder.py:62-66,137-139): geometry as STORE_CONVT, block_n == ctCout;
   * out_f32[g, c, t', h', w'] = sum_co act(acc+bias)[tap, co] * w2[g, c, co], g = row / rows_per_group, c < c2 <= 4 */
  int64_t rows_per_group;
  int block_n;         /* N tile (multiple of 16, <= 256); 0 = choose                              */
  int cta_pair;        /* 0 = auto, 1 = force the 2-CTA (cta_group::2, 256-row tile) kernel, -1 = never */
  void* prof;          /* NULL, or device int64[3*512]: clock64 timeline of CTA 0 (producer / MMA / epilogue; tuning aid) */
  /* Split-K (ROWMAJOR only): problems with few output tiles and a long K loop (low-resolution conv pyramid levels)
   * are cut along K into split_k slices whose fp32 partial sums are atomically accumulated into splitk_ws, followed by
   * a finalize kernel (bias / activation / residual / stores) that also re-zeroes the workspace.
   * splitk_ws: NULL (never split) or a ZERO-FILLED device buffer of splitk_ws_bytes >= M*N*4 that no concurrently
   * running l4p_gemm uses; split_k: 0 = choose, 1 = off, >1 = forced slice count. */
  void* splitk_ws;
  int64_t splitk_ws_bytes;
  int split_k;
  /* Grouped weights (A_MATRIX + ROWMAJOR, 1-CTA kernel): the rows of A form groups of grp_a_rows consecutive rows, group g
   * multiplies its OWN weight block: W rows [g*grp_b_rows, g*grp_b_rows + N) of a [groups*grp_b_rows, ldw] matrix
   * (grp_a_rows = 0: one shared W). m_stride (0 = 128): consecutive 128-row tiles start m_stride rows apart and only the
   * first m_stride rows of a tile are stored, so that groups smaller than a tile (grp_a_rows == m_stride < 128) still get
   * one tile each. Used by the track head's token -> video-token attention: per-query score matrices
   * S_g^T = Q'_g X_g^T (grp_a_rows = m_stride = 48 head-token rows, W = the query's 2048 video tokens) and the per-query
   * output projection of the video-token -> token attention (sam/transformer.py:223-245 applied through
   * (q W_k^T) x instead of q (W_k x)). */
  int64_t grp_a_rows, grp_b_rows;
  int m_stride;
  /* A_CONV3D with grouped weights (ROWMAJOR store): the batch axis holds conv_grp_b entries per group (cB = groups *
   * conv_grp_b); group g convolves with W rows [g*N, (g+1)*N) of a [groups*N, K] weight stack and adds bias[g*N ...]:
   * the identical layers of several decoder heads (DPT flow / depth / motion-mask heads, dpt_block.py) in ONE launch. */
  int conv_grp_b;
} l4p_gemm_desc;

/* Replaces F.linear/addmm (modeling_finetune.py:62-69,171-177,188), the 1x1x1/3x3x3 Conv3d and k==s
 * ConvTranspose3d calls of the DPT heads (task_heads/dpt/croco/dpt_block.py:29-90,144-157,255-278,406-414)
 * and the SAM projections (task_heads/sam/transformer.py:223-245). */
int l4p_gemm(const l4p_gemm_desc* desc, void* stream);

/* Host-only dry run of l4p_gemm: validates the descriptor and reports the launch decisions without touching the device
 * (pointers in `desc` are not dereferenced). out6 = {block_n, split_k, cta_pair (0/1), smem ring stages, grid CTAs,
 * threads per CTA}. Usable without a GPU (the SM count then defaults to 148). */
int l4p_gemm_plan(const l4p_gemm_desc* desc, int* out6);

/* ---- K1/K7/K9 helpers (HBM-bound) ------------------------------------------------------------- */
/* Tubelet gather for the patch embedding: rgb fp32 [B,C,T,H,W] -> 16-bit [B*(T/pt)*(H/ph)*(W/pw), C*pt*ph*pw],
 * K ordered (c,dt,dh,dw) = flattened Conv3d weight; token order t'*nh*nw + h'*nw + w'.
 * With l4p_gemm this replaces PatchEmbed.forward (modeling_finetune.py:276-283). */
int l4p_patchify(const float* rgb, void* out16, int B, int C, int T, int H, int W, int pt, int ph, int pw,
                 int bf16, void* stream);

/* N3 video preprocessing (l4p/data/l4p_dataset_mini.py:543-587, rgb key): temporal mirror padding (:126-190) ->
 * bilinear spatial resize of the uint8 frames [T0,H0,W0,3] to (Hs,Ws) (F.interpolate 'trilinear' with T unchanged,
 * align_corners=False, :237-290) -> crop of (To,Hc,Wc) at (t0,i0,j0) (:292-345) -> (x/255 - mean)/std (:576-580),
 * fused into one gather that writes one clip of rgb_b3thw: out fp32 [3,To,Hc,Wc]. mean3/std3 are HOST pointers. */
int l4p_preprocess_rgb(const void* frames_u8, float* out, int T0, int H0, int W0, int Hs, int Ws, int t0, int i0, int j0,
                       int To, int Hc, int Wc, const float* mean3, const float* std3, void* stream);
/* fp32 -> 16-bit cast, n a multiple of 4. */
int l4p_cast16(const float* x, void* y16, int64_t n, int bf16, void* stream);
/* The same with an optional fp32 addend of the same length (y = round16(x + add); add may be NULL): the "+ positional
 * embedding" of the SAM attention inputs (sam/transformer.py:168-170,178-180) folded into the operand cast. */
int l4p_cast16_add(const float* x, const float* add, void* y16, int64_t n, int bf16, void* stream);
/* Trilinear resampling of channels-last 16-bit [B,Ti,Hi,Wi,C] -> [B,To,Ho,Wo,C] (y16 and/or its ReLU y16_relu).
 * Replaces F.interpolate(mode="trilinear") (dpt_block.py:231-236, dpt_head.py:79-83: align_corners=1;
 * sparse_heads.py:645-647: align_corners=0). */
int l4p_upsample3d(const void* x16, void* y16, void* y16_relu, int B, int Ti, int Hi, int Wi, int To, int Ho,
                   int Wo, int C, int align_corners, int bf16, void* stream);
/* 3x3x3 / pad 1 / stride (sT,sH,sW) gather of channels-last x into [B*To*Ho*Wo, 27*C] rows (kt,kh,kw,c);
 * with l4p_gemm this replaces the stride-2 Conv3d of the DPT reassemble stage (dpt_block.py:265-278). */
int l4p_im2col3(const void* x16, void* out16, int B, int T, int H, int W, int C, int sT, int sH, int sW,
                void* stream);

/* ---- K13-K15: track head (SAM two-way transformer + mask decoder read-outs) -------------------- */
/* Few queries x many keys: out[g,j,h*d+c] = softmax_k(scale * q[g,j,h,:].k[g,k,h,:]) v[g,k,h,c].
 * q,out fp32 [G,nq<=8,H*d]; k16,v16 16-bit rows [.., H*d], group g starts at row g*kv_group_rows (0 = shared).
 * Replaces Attention.forward for token->image and token self attention (sam/transformer.py:223-245). */
int l4p_token_attention(const float* q, const void* k16, const void* v16, float* out, int G, int nq, int Nk,
                        int H, int d, int64_t kv_group_rows, float scale, int bf16, void* stream);
/* Many queries x few keys (image -> token cross attention, sam/transformer.py:179-184):
 * q16,out16 16-bit [G*Np, H*d]; k,v fp32 [G,nk<=8,H*d]. This build: d == 88. */
int l4p_image_attention(const void* q16, const float* k, const float* v, void* out16, int G, int Np, int nk,
                        int H, int d, float scale, int bf16, void* stream);

/* LayerNorm over the channels of 16-bit rows (+GELU): LayerNorm3d + activation of MaskDecoder.output_upscaling
 * (sam/mask_decoder.py:58-66,145-157). */
int l4p_layernorm16(const void* x16, const float* gamma, const float* beta, void* y16, int64_t rows, int cols,
                    float eps, int gelu, int bf16, void* stream);
/* Token -> video-token attention with the projections moved to the token side (sam/transformer.py:223-245, see
 * csrc/track_t2i.cu): score = x . (W_k^T q) + q . (W_k pe + b_k), out = W_v (sum_n p_n x_n) + b_v.
 * l4p_head_expand: q fp32 [G*nt, heads*hd] -> out16 [G*heads*nt, heads*hd], row (g, h, t) = q[g,t] * scale restricted to the
 * columns of head h (zero elsewhere): the block-diagonal operand that turns per-head dot products into plain GEMMs. */
int l4p_head_expand(const float* q, void* out16, int64_t G, int nt, int heads, int hd, float scale, int bf16, void* stream);
/* Inverse bookkeeping: out fp32 [G*nt, heads*hd], out[(g,t), h*hd + d] = z[(g,h,t), h*hd + d] for z fp32 [G*heads*nt, heads*hd]. */
int l4p_head_diag_gather(const float* z, float* out, int64_t G, int nt, int heads, int hd, void* stream);
/* Row softmax of fp32 scores [rows, n] (already scaled) -> 16-bit probabilities; n multiple of 4, <= 2048. */
int l4p_row_softmax16(const float* s, void* p16, int64_t rows, int n, int bf16, void* stream);
/* Video-token -> token attention, folded form: s fp32 [G*heads*nt, n] (transposed scores, row (g,h,t)) -> softmax over the nt
 * tokens of each head -> p16 [G*n, heads*nt] (the A operand of the per-query K = heads*nt output GEMM); nt <= 8. */
int l4p_group_softmax_t16(const float* s, void* p16, int64_t G, int heads, int nt, int n, int bf16, void* stream);
/* y16[g] (J x C) = p16[g] (J x n) * x16[g] (n x C) for G groups: p16 [G*J, n], x16 [G*n, C], y16 [G*J, C]; fp32 accumulation,
 * J <= 48, n multiple of 64, C multiple of 16. */
int l4p_token_weighted_sum(const void* p16, const void* x16, void* y16, int64_t G, int J, int n, int C, int bf16, void* stream);

/* masks fp32 [G,nch<=3,T,h,w] low-res logits -> bilinear (align_corners=False) to (H,W) fused with the
 * read-outs: traj[G,2,T] = soft-argmax of channel 0 at pixel centres (+0.5), vis[G,1,T] = mean of channel 1,
 * depth[G,1,T] = exp(mean of channel 2) (sparse_heads.py:140-160,574-589,645-647). */
int l4p_track_readout(const float* masks, float* traj, float* vis, float* depth, int G, int nch, int T, int h,
                      int w, int H, int W, void* stream);

/* ---- K4: fused attention -------------------------------------------------------------------- */
/* out[b*N+n, h*head_dim + c] = sum_m softmax_m(scale * q[b,h,n,:] . k[b,h,m,:]) * v[b,h,m,c]
 * Replaces the q@k^T -> softmax -> @v sequence of Attention.forward
 * (l4p/models/VideoMAEv2/models/modeling_finetune.py:180-186); the [B,H,N,N] score tensor is never
 * materialised. q,k: [B,H,N,head_dim_pad]; vt: [B,H,head_dim_pad,N] (pad lanes zero); out 16-bit
 * [B*N, H*head_dim]. This build: head_dim_pad == 96, N a multiple of 256.
 * prof: NULL, or a device buffer of 3*64*8 + 5*grid int64 that receives clock64 stamps of CTA 0 followed by
 * per-CTA {globaltimer start, end, clock64 start, end, smid} (kernel tuning aid). */
int l4p_attention(const void* q, const void* k, const void* vt, void* out, int B, int H, int N,
                  int head_dim, int head_dim_pad, float scale, int bf16, void* stream, void* prof);

/* ---- K11: camera pose from Plücker rays ------------------------------------------------------ */
/* rays fp32 [B,6,T,h,w] (direction, moment). mode 0: use normalised intrinsics k_norm [B,4,4,T]
 * (rays_to_cameras, l4p/utils/geometry_utils.py:331-406); mode 1: estimate one fixed K from frame 0 by a
 * normalised-DLT homography with `refits` consensus refits at `reproj_threshold`, RQ-decomposed
 * (rays_to_cameras_and_fixed_per_frame_intrinsics, :493-579, replacing cv2.findHomography/RQDecomp3x3 :436-448),
 * k_est [B,4,4,T] is reported at (outH,outW). Outputs: ext [B,4,4,T] (cam<-world), pose [B,4,4,T] = ext^-1
 * (dense_heads.py:346), centers [B,T,3]. ws_kgrid: B*9 doubles of scratch (mode 1). */
int l4p_pose_from_rays(const float* rays, const float* k_norm, int mode, int B, int T, int h, int w, int outH,
                       int outW, float reproj_threshold, int refits, double* ws_kgrid, float* ext, float* pose,
                       float* centers, float* k_est, void* stream);

/* ---- K16: depth window affine aligner ----------------------------------------------------------- */
/* sol[b] = argmin_(s,t) || s*f(pred[b]) + t - f(target[b]) ||^2 over n elements, f = safe_inverse when
 * inverse != 0 (LstSqAffineAligner.solve, l4p/models/aligner.py:45-57; misc.py:48-62). ws_moments: 5*B doubles. */
int l4p_affine_align_solve(const float* pred, const float* target, int B, int64_t n, int64_t pred_stride,
                           int64_t target_stride, int inverse, double* ws_moments, float* sol, void* stream);
/* y = f(s*f(x)+t) (LstSqAffineAligner.apply, aligner.py:59-66); x,y [B,n]. */
int l4p_affine_align_apply(const float* x, float* y, const float* sol, int B, int64_t n, int inverse, void* stream);

/* ---- K17: joint depth + pose similarity aligner --------------------------------------------------- */
/* Similarity transform dst ~ s R src + t between the point maps of the overlap frames (every frame_step-th of
 * `overlap`) of the current window (src) and of the stitched buffer (dst): weighted Umeyama iterated `iters`
 * times with consensus re-selection at *thr_dev (device scalar). Replaces KabaschUmeyama3DAligner.solve
 * (l4p/models/aligner.py:177-237: CPU numpy + skimage RANSAC) and generate_point_map (geometry_utils.py:13-53).
 * depth_* fp32 [T_*,H,W] (one clip), K_* fp32 [4,4,T_*] pixel intrinsics, pose_* fp32 [16,T_*] (cam->world).
 * Result in ws[17..33]: 4x4 row-major T (doubles) followed by the scale; ws[0..16] is scratch. */
int l4p_sim3_align(const float* depth_src, const float* K_src, const float* pose_src, int T_src,
                   const float* depth_dst, const float* K_dst, const float* pose_dst, int T_dst, int overlap,
                   int frame_step, int H, int W, const float* thr_dev, int iters, int min_points, double* ws,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* L4P_B200_H_ */
