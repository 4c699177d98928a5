// Track-head kernels (K13-K15): the token<->image attentions of SAM's two-way transformer are extremely
// skinny (6 prompt tokens per query against 2048 video tokens), so they run on CUDA cores and are HBM-bound
// on the per-query K/V/Q projections; the big projections themselves go through the tcgen05 GEMM (gemm.cu).
//
//   l4p_token_attention   few queries x many keys   (sam/transformer.py:223-245 as used at :163-171,104-107,
//                         and the 6x6 token self-attention :157-161)
//   l4p_image_attention   many queries x few keys   (image -> token cross attention, :179-184)
//   l4p_layernorm16       channel LayerNorm on 16-bit rows with optional GELU (LayerNorm3d + GELU of the mask
//                         decoder's upscaling path, sam/mask_decoder.py:58-66,145-157)
//   l4p_track_readout     fused trilinear upsample (align_corners=False) of the low-res mask logits to the image
//                         size + soft-argmax / visibility mean / depth exp-mean (sparse_heads.py:140-160,574-589,
//                         645-647): the [Nq,3,16,224,224] logits (1.2 GB at Nq=128) never exist in HBM.
#include "common.cuh"

namespace l4p {

constexpr int kTokMaxQ = 8;     // prompt tokens per group (6 used)
constexpr int kTokMaxPairs = 3; // channel pairs per lane: head_dim <= 192

// ------------------------------------------------------------------------------------------------
// q fp32 [G, nq, ldq] ; k16/v16 rows [.., ldkv] 16-bit, group g starts at row g*kv_group_rows (0 = shared) ;
// out fp32 [G, nq, ldq]. grid (H, G), 256 threads, dynamic smem: nq*Nk floats (scores) + nq*d (q) + 8*nq*d (partials)
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256)
token_attention_kernel(const float* __restrict__ q, const uint16_t* __restrict__ k16, const uint16_t* __restrict__ v16,
                       float* __restrict__ out, int nq, int Nk, int d, int ldq, int ldkv, long long kv_group_rows,
                       float scale) {
  extern __shared__ float sm[];
  float* s_scores = sm;                    // [nq][Nk]
  float* s_q = s_scores + nq * Nk;         // [nq][d]
  float* s_part = s_q + nq * d;            // [8][nq][d]
  __shared__ float s_max[kTokMaxQ], s_sum[kTokMaxQ];
  const int h = blockIdx.x, g = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* qg = q + ((long long)g * nq) * ldq + h * d;
  const uint16_t* kg = k16 + (long long)g * kv_group_rows * ldkv + h * d;
  const uint16_t* vg = v16 + (long long)g * kv_group_rows * ldkv + h * d;
  for (int i = tid; i < nq * d; i += blockDim.x) s_q[i] = qg[(i / d) * ldq + (i % d)] * scale;
  __syncthreads();
  // pass 1: one key per thread, d multiple of 8 (16-byte chunks)
  for (int key = tid; key < Nk; key += blockDim.x) {
    float acc[kTokMaxQ];
#pragma unroll
    for (int j = 0; j < kTokMaxQ; ++j) acc[j] = 0.f;
    const uint4* kr = reinterpret_cast<const uint4*>(kg + (long long)key * ldkv);
    for (int c8 = 0; c8 < d / 8; ++c8) {
      const uint4 raw = kr[c8];
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
      float kv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack2<BF16>(w[i]);
        kv[2 * i] = f.x; kv[2 * i + 1] = f.y;
      }
#pragma unroll
      for (int j = 0; j < kTokMaxQ; ++j) {
        if (j < nq) {
          const float* qq = s_q + j * d + c8 * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j] = fmaf(kv[i], qq[i], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kTokMaxQ; ++j)
      if (j < nq) s_scores[j * Nk + key] = acc[j];
  }
  __syncthreads();
  // softmax statistics: warp j handles query row j
  if (warp < nq) {
    float m = -INFINITY;
    for (int key = lane; key < Nk; key += 32) m = fmaxf(m, s_scores[warp * Nk + key]);
    m = warp_max(m);
    float s = 0.f;
    for (int key = lane; key < Nk; key += 32) {
      const float p = __expf(s_scores[warp * Nk + key] - m);
      s_scores[warp * Nk + key] = p;
      s += p;
    }
    s = warp_sum(s);
    if (lane == 0) { s_max[warp] = m; s_sum[warp] = s; }
  }
  __syncthreads();
  // pass 2: each warp takes a strided subset of keys, lanes cover channel pairs
  float o[kTokMaxQ][kTokMaxPairs][2];
#pragma unroll
  for (int j = 0; j < kTokMaxQ; ++j)
#pragma unroll
    for (int c = 0; c < kTokMaxPairs; ++c) o[j][c][0] = o[j][c][1] = 0.f;
  const int npairs = d / 2;
  for (int key = warp; key < Nk; key += 8) {
    const uint32_t* vr = reinterpret_cast<const uint32_t*>(vg + (long long)key * ldkv);
#pragma unroll
    for (int c = 0; c < kTokMaxPairs; ++c) {
      const int cp = lane + c * 32;
      if (cp < npairs) {
        const float2 f = unpack2<BF16>(vr[cp]);
#pragma unroll
        for (int j = 0; j < kTokMaxQ; ++j) {
          if (j < nq) {
            const float p = s_scores[j * Nk + key];
            o[j][c][0] = fmaf(p, f.x, o[j][c][0]);
            o[j][c][1] = fmaf(p, f.y, o[j][c][1]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kTokMaxQ; ++j)
    if (j < nq)
#pragma unroll
      for (int c = 0; c < kTokMaxPairs; ++c) {
        const int cp = lane + c * 32;
        if (cp < npairs) {
          s_part[(warp * nq + j) * d + 2 * cp] = o[j][c][0];
          s_part[(warp * nq + j) * d + 2 * cp + 1] = o[j][c][1];
        }
      }
  __syncthreads();
  for (int i = tid; i < nq * d; i += blockDim.x) {
    const int j = i / d;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += s_part[(w * nq + j) * d + (i % d)];
    out[((long long)g * nq + j) * ldq + h * d + (i % d)] = a / s_sum[j];
  }
}

// ------------------------------------------------------------------------------------------------
// Hot-shape specialisation of token_attention: head_dim D compile-time (88), nq <= 6, many keys.
//   pass 1  thread-per-key, channel-chunk outer loop: the 6 x 8 query values of a chunk sit in registers and are
//           reused for the thread's KPT keys; every 16-byte key load is independent (deep memory-level parallelism)
//   pass 2  half-warps walk the keys, 11 lanes x 16 bytes cover one V row, 48 accumulators per lane
// ------------------------------------------------------------------------------------------------
template <bool BF16, int D, int NQ>
__global__ void __launch_bounds__(256, 2)
token_attention_fast_kernel(const float* __restrict__ q, const uint16_t* __restrict__ k16, const uint16_t* __restrict__ v16,
                            float* __restrict__ out, int Nk, int ldq, int ldkv, long long kv_group_rows, float scale) {
  constexpr int C8 = D / 8;       // 16-byte chunks per row
  constexpr int KPT = 4;          // keys per thread per pass-1 sweep (ncu: KPT = 8 needed 184 registers -> 1 block / SM, 12 % occupancy)
  extern __shared__ float sm[];
  float* s_scores = sm;                 // [NQ][Nk]
  float* s_q = s_scores + NQ * Nk;      // [NQ][D]
  float* s_part = s_q + NQ * D;         // [16][NQ][D]
  __shared__ float s_sum[NQ];
  const int h = blockIdx.x, g = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* qg = q + ((long long)g * NQ) * ldq + h * D;
  const uint16_t* kg = k16 + (long long)g * kv_group_rows * ldkv + h * D;
  const uint16_t* vg = v16 + (long long)g * kv_group_rows * ldkv + h * D;
  for (int i = tid; i < NQ * D; i += 256) s_q[i] = qg[(i / D) * ldq + (i % D)] * scale;
  __syncthreads();
  for (int kbase = 0; kbase < Nk; kbase += 256 * KPT) {
    float acc[KPT][NQ];
#pragma unroll
    for (int kk = 0; kk < KPT; ++kk)
#pragma unroll
      for (int j = 0; j < NQ; ++j) acc[kk][j] = 0.f;
#pragma unroll 1
    for (int c8 = 0; c8 < C8; ++c8) {
      float qq[NQ][8];
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(s_q + j * D + c8 * 8);
        const float4 b = *reinterpret_cast<const float4*>(s_q + j * D + c8 * 8 + 4);
        qq[j][0] = a.x; qq[j][1] = a.y; qq[j][2] = a.z; qq[j][3] = a.w;
        qq[j][4] = b.x; qq[j][5] = b.y; qq[j][6] = b.z; qq[j][7] = b.w;
      }
      uint4 raw[KPT];
#pragma unroll
      for (int kk = 0; kk < KPT; ++kk) {
        const int key = kbase + tid + kk * 256;
        raw[kk] = key < Nk ? *reinterpret_cast<const uint4*>(kg + (long long)key * ldkv + c8 * 8) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int kk = 0; kk < KPT; ++kk) {
        const uint32_t w[4] = {raw[kk].x, raw[kk].y, raw[kk].z, raw[kk].w};
        float kv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack2<BF16>(w[i]);
          kv[2 * i] = f.x; kv[2 * i + 1] = f.y;
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[kk][j] = fmaf(kv[i], qq[j][i], acc[kk][j]);
      }
    }
#pragma unroll
    for (int kk = 0; kk < KPT; ++kk) {
      const int key = kbase + tid + kk * 256;
      if (key < Nk)
#pragma unroll
        for (int j = 0; j < NQ; ++j) s_scores[j * Nk + key] = acc[kk][j];
    }
  }
  __syncthreads();
  if (warp < NQ) {
    float m = -INFINITY;
    for (int key = lane; key < Nk; key += 32) m = fmaxf(m, s_scores[warp * Nk + key]);
    m = warp_max(m);
    float sum = 0.f;
    for (int key = lane; key < Nk; key += 32) {
      const float pv = __expf(s_scores[warp * Nk + key] - m);
      s_scores[warp * Nk + key] = pv;
      sum += pv;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_sum[warp] = sum;
  }
  __syncthreads();
  {
    const int half = lane >> 4, l16 = lane & 15;  // two keys per warp iteration; lanes 0..C8-1 of each half active
    float o[NQ][8];
#pragma unroll
    for (int j = 0; j < NQ; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) o[j][i] = 0.f;
    if (l16 < C8) {
#pragma unroll 4
      for (int key = warp * 2 + half; key < Nk; key += 16) {
        const uint4 raw = *reinterpret_cast<const uint4*>(vg + (long long)key * ldkv + l16 * 8);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        float vv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack2<BF16>(w[i]);
          vv[2 * i] = f.x; vv[2 * i + 1] = f.y;
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          const float pj = s_scores[j * Nk + key];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[j][i] = fmaf(pj, vv[i], o[j][i]);
        }
      }
#pragma unroll
      for (int j = 0; j < NQ; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) s_part[((warp * 2 + half) * NQ + j) * D + l16 * 8 + i] = o[j][i];
    }
  }
  __syncthreads();
  for (int i = tid; i < NQ * D; i += 256) {
    const int j = i / D;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) a += s_part[(w * NQ + j) * D + (i % D)];
    out[((long long)g * NQ + j) * ldq + h * D + (i % D)] = a / s_sum[j];
  }
}

// ------------------------------------------------------------------------------------------------
// q16 [G*Np, H*d] ; k,v fp32 [G, nk, H*d] ; out16 [G*Np, H*d]. One thread per (row, head): warp = head, lane = row, so
// that every shared-memory read of k / v is a warp-wide broadcast (one wavefront) and the 2 * nk * d FMAs per thread
// run two-wide on the packed f32x2 pipe. blockDim = 32 * H.
// ------------------------------------------------------------------------------------------------
// Rows are staged through shared memory: a block of 32 * H threads loads 32 rows (32 x H*D 16-bit values) with fully
// coalesced 16-byte accesses, every thread then works on its own (row, head) slice out of shared memory (row pitch padded by
// 16 bytes: conflict-free 128-bit reads for lane = row), overwrites the slice with its output, and the tile is written back
// coalesced. ncu on the direct version (lane = row reading global memory): every 16-byte access touched 32 different lines,
// 1.5 TB/s at ideal DRAM traffic.
template <bool BF16, int D>
__global__ void __launch_bounds__(256, 2)
image_attention_kernel(const uint16_t* __restrict__ q16, const float* __restrict__ kf, const float* __restrict__ vf,
                       uint16_t* __restrict__ out16, int Np, int nk, int H, float scale, int rows_per_block) {
  extern __shared__ float sm[];  // k [nk][H*D] (pre-scaled), v [nk][H*D], then the 16-bit row tile [32][H*D + 8]
  const int ld = H * D;
  const int pitch = ld + 8;      // 16-bit elements
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  const int g = (int)(row0 / Np);
  float* s_k = sm;
  float* s_v = sm + nk * ld;
  uint16_t* s_q = reinterpret_cast<uint16_t*>(sm + 2 * nk * ld);
  for (int i = threadIdx.x; i < nk * ld; i += blockDim.x) {
    s_k[i] = kf[(long long)g * nk * ld + i] * scale;
    s_v[i] = vf[(long long)g * nk * ld + i];
  }
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vec_per_row = ld / 8;  // 16-byte vectors per row
  for (int r0 = 0; r0 < rows_per_block; r0 += 32) {
    __syncthreads();  // k/v ready (first tile); previous tile fully written back
    for (int i = threadIdx.x; i < 32 * vec_per_row; i += blockDim.x) {
      const int rr = i / vec_per_row, cv = i - rr * vec_per_row;
      *reinterpret_cast<uint4*>(s_q + rr * pitch + cv * 8) =
          *reinterpret_cast<const uint4*>(q16 + (row0 + r0 + rr) * ld + cv * 8);
    }
    __syncthreads();
    uint16_t* mine = s_q + lane * pitch + h * D;
    float sc[kTokMaxQ];
#pragma unroll
    for (int j = 0; j < kTokMaxQ; ++j) sc[j] = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < D / 8; ++c8) {
      const uint4 raw = *reinterpret_cast<const uint4*>(mine + c8 * 8);
      const float2 q0 = unpack2<BF16>(raw.x), q1 = unpack2<BF16>(raw.y), q2 = unpack2<BF16>(raw.z), q3 = unpack2<BF16>(raw.w);
#pragma unroll
      for (int j = 0; j < kTokMaxQ; ++j) {
        if (j < nk) {
          const float4* kk = reinterpret_cast<const float4*>(s_k + j * ld + h * D + c8 * 8);
          const float4 ka = kk[0], kb = kk[1];
          float a = sc[j];
          a = fmaf(q0.x, ka.x, a); a = fmaf(q0.y, ka.y, a); a = fmaf(q1.x, ka.z, a); a = fmaf(q1.y, ka.w, a);
          a = fmaf(q2.x, kb.x, a); a = fmaf(q2.y, kb.y, a); a = fmaf(q3.x, kb.z, a); a = fmaf(q3.y, kb.w, a);
          sc[j] = a;
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kTokMaxQ; ++j)
      if (j < nk) mx = fmaxf(mx, sc[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kTokMaxQ; ++j) {
      sc[j] = j < nk ? __expf(sc[j] - mx) : 0.f;
      sum += sc[j];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < kTokMaxQ; ++j) sc[j] *= inv;
#pragma unroll
    for (int c8 = 0; c8 < D / 8; ++c8) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
      for (int j = 0; j < kTokMaxQ; ++j) {
        if (j < nk) {
          const float4* vv = reinterpret_cast<const float4*>(s_v + j * ld + h * D + c8 * 8);
          const float4 va = vv[0], vb = vv[1];
          const float pj = sc[j];
          o[0] = fmaf(pj, va.x, o[0]); o[1] = fmaf(pj, va.y, o[1]); o[2] = fmaf(pj, va.z, o[2]); o[3] = fmaf(pj, va.w, o[3]);
          o[4] = fmaf(pj, vb.x, o[4]); o[5] = fmaf(pj, vb.y, o[5]); o[6] = fmaf(pj, vb.z, o[6]); o[7] = fmaf(pj, vb.w, o[7]);
        }
      }
      // the slice is private to this thread and fully consumed by the score pass: overwrite it with the output
      *reinterpret_cast<uint4*>(mine + c8 * 8) =
          make_uint4(pack2<BF16>(o[0], o[1]), pack2<BF16>(o[2], o[3]), pack2<BF16>(o[4], o[5]), pack2<BF16>(o[6], o[7]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * vec_per_row; i += blockDim.x) {
      const int rr = i / vec_per_row, cv = i - rr * vec_per_row;
      *reinterpret_cast<uint4*>(out16 + (row0 + r0 + rr) * ld + cv * 8) = *reinterpret_cast<const uint4*>(s_q + rr * pitch + cv * 8);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over 16-bit rows (+ optional GELU). LPR lanes cooperate on one row (32/LPR rows per warp in flight,
// which is what makes short rows like the 352-channel LayerNorm3d bandwidth- instead of latency-bound);
// cols multiple of 8, cols <= LPR * 8 * kLn16Iters.
// ------------------------------------------------------------------------------------------------
constexpr int kLn16Iters = 8;
template <bool BF16, int LPR, int ITERS>  // cols <= LPR * 8 * ITERS; fewer iterations = fewer registers = more rows in flight
__global__ void __launch_bounds__(256, ITERS <= 6 ? 3 : 2)
layernorm16_kernel(const uint16_t* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   uint16_t* __restrict__ y, long long rows, int cols, float eps, int gelu) {
  constexpr int RPW = 32 / LPR;  // rows per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * RPW + sub;
  const bool ok = row < rows;
  const int nch = cols >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (ok ? row : 0) * cols);
  float v[ITERS][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int j = l + i * LPR;
    if (ok && j < nch) {
      const uint4 raw = xr[j];
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = unpack2<BF16>(w[t]);
        v[i][2 * t] = f.x; v[i][2 * t + 1] = f.y;
        s += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)cols;
  float qq = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int j = l + i * LPR;
    if (ok && j < nch)
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float dlt = v[i][t] - mean; qq += dlt * dlt; }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
  const float rstd = rsqrtf(qq / (float)cols + eps);
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int j = l + i * LPR;
    if (ok && j < nch) {
      const float4 g0 = reinterpret_cast<const float4*>(gamma)[2 * j], g1 = reinterpret_cast<const float4*>(gamma)[2 * j + 1];
      const float4 b0 = reinterpret_cast<const float4*>(beta)[2 * j], b1 = reinterpret_cast<const float4*>(beta)[2 * j + 1];
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] = (v[i][t] - mean) * rstd * gg[t] + bb[t];
      if (gelu) {
#pragma unroll
        for (int t = 0; t < 8; t += 2) gelu2(o[t], o[t + 1]);
      }
      reinterpret_cast<uint4*>(y + row * cols)[j] = make_uint4(pack2<BF16>(o[0], o[1]), pack2<BF16>(o[2], o[3]),
                                                               pack2<BF16>(o[4], o[5]), pack2<BF16>(o[6], o[7]));
    }
  }
}

// Long rows (the 1408-channel per-query token stream of the two-way layers, 16-bit in and out): one warp owns R rows at a
// time and keeps them as RAW packed words (CH x uint4 per lane and row instead of 8 x CH floats), so all R x CH 16-byte
// loads of a lane are in flight at once: 11 KB per warp, ~135 KB per SM (R = 4, 12 warps per SM at <= 168 registers). The
// generic kernel above holds one unpacked row per warp (2.8 KB in flight) and reached 3.1 TB/s on [262144, 1408].
// With the rows in flight the kernel is INSTRUCTION-bound (16-bit input doubles the elements per byte): statistics in ONE
// pass as shifted moments (shift = the row's first element, so E[(x-s)^2] - E[x-s]^2 does not cancel) and all arithmetic
// on the packed f32x2 pipe: ~5.5 instead of ~9.5 instructions per element.
template <bool BF16, int CH, int R>  // cols <= 32 * 8 * CH
__global__ void __launch_bounds__(128, 3)
layernorm16_rows_kernel(const uint16_t* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                        uint16_t* __restrict__ y, long long rows, int cols, float eps, int gelu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * R;
  const int nch = cols >> 3;
  uint4 raw[R][CH];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool ok = row0 + r < rows;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (ok ? row0 + r : 0) * cols);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int j = lane + i * 32;
      raw[r][i] = (ok && j < nch) ? xr[j] : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  float mean[R], rstd[R];
  {
    uint64_t s1[R], s2[R];
    float shift[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      shift[r] = __shfl_sync(0xffffffffu, unpack2<BF16>(raw[r][0].x).x, 0);   // the row's first element
      const uint64_t ns2 = pk2(-shift[r], -shift[r]);
      uint64_t a1 = pk2(0.f, 0.f), a2 = pk2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (lane + i * 32 < nch) {
          const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack2<BF16>(w[t]);
            const uint64_t d = add2(pk2(f.x, f.y), ns2);
            a1 = add2(a1, d);
            a2 = fma2(d, d, a2);
          }
        }
      }
      s1[r] = a1; s2[r] = a2;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float u0, u1, q0, q1;
      upk2(s1[r], u0, u1);
      upk2(s2[r], q0, q1);
      float u = u0 + u1, q = q0 + q1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        u += __shfl_xor_sync(0xffffffffu, u, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      const float m = u / (float)cols;
      mean[r] = shift[r] + m;
      rstd[r] = rsqrtf(fmaxf(q / (float)cols - m * m, 0.f) + eps);
    }
    // launder the raw words: without this the compiler keeps the UNPACKED values of the statistics pass alive for the
    // output pass (common-subexpression elimination), i.e. 8 x CH x R floats, and spills
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < CH; ++i)
        asm volatile("" : "+r"(raw[r][i].x), "+r"(raw[r][i].y), "+r"(raw[r][i].z), "+r"(raw[r][i].w));
  }
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int j = lane + i * 32;
    asm volatile("" ::: "memory");  // keep the gamma / beta loads of later chunks from being hoisted (they would spill the rows)
    if (j < nch) {
      const float4 g0 = reinterpret_cast<const float4*>(gamma)[2 * j], g1 = reinterpret_cast<const float4*>(gamma)[2 * j + 1];
      const float4 b0 = reinterpret_cast<const float4*>(beta)[2 * j], b1 = reinterpret_cast<const float4*>(beta)[2 * j + 1];
      const uint64_t gg[4] = {pk2(g0.x, g0.y), pk2(g0.z, g0.w), pk2(g1.x, g1.y), pk2(g1.z, g1.w)};
      const uint64_t bb[4] = {pk2(b0.x, b0.y), pk2(b0.z, b0.w), pk2(b1.x, b1.y), pk2(b1.z, b1.w)};
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (row0 + r < rows) {
          const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
          const uint64_t rs2 = pk2(rstd[r], rstd[r]), nm2 = pk2(-mean[r], -mean[r]);
          float o[8];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = unpack2<BF16>(w[t]);
            // ((x - mean) * rstd) * gamma + beta, the association of the other LayerNorm kernels
            const uint64_t v = fma2(mul2(add2(pk2(f.x, f.y), nm2), rs2), gg[t], bb[t]);
            upk2(v, o[2 * t], o[2 * t + 1]);
          }
          if (gelu) {
#pragma unroll
            for (int t = 0; t < 8; t += 2) gelu2(o[t], o[t + 1]);
          }
          reinterpret_cast<uint4*>(y + (row0 + r) * cols)[j] = make_uint4(pack2<BF16>(o[0], o[1]), pack2<BF16>(o[2], o[3]),
                                                                          pack2<BF16>(o[4], o[5]), pack2<BF16>(o[6], o[7]));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// masks fp32 [G, 3, T, h, w] (low-res logits) -> traj [G,2,T], vis [G,1,T], depth [G,1,T]. One block per (g,t).
// Spatial bilinear upsample to (H,W) with align_corners=False; T is not resampled (T_in == T_out).
// ------------------------------------------------------------------------------------------------
// align_corners=False source coordinate of output index o: (i0, i1, fraction)
L4P_DEVICE void src_coord(int o, int n_in, int n_out, int& i0, int& i1, float& f) {
  float sc = ((float)o + 0.5f) * ((float)n_in / (float)n_out) - 0.5f;
  sc = sc < 0.f ? 0.f : sc;
  i0 = (int)sc;
  i1 = i0 + 1 < n_in ? i0 + 1 : n_in - 1;
  f = sc - (float)i0;
}

// One block per (g,t). The upsampled logits are never formed for the two MEAN channels (visibility, depth): the mean
// of a bilinear upsample is a separable weighted sum of the h x w source (weights = how much every source row /
// column contributes to all output rows / columns). The soft-argmax channel evaluates the H x W upsample row by row
// (row interpolation hoisted, per-column (x0, x1, fx) from a shared table) with ONE exp per pixel: the source maximum
// bounds every interpolated value, so it is a valid softmax stabiliser and no online rescaling is needed.
__global__ void __launch_bounds__(256)
track_readout_kernel(const float* __restrict__ masks, float* __restrict__ traj, float* __restrict__ vis,
                     float* __restrict__ depth, int T, int h, int w, int H, int W, int has_vis, int has_depth,
                     int nch) {
  extern __shared__ float sm[];  // src [h*w] | wy [h] | wx [w] | fx [W] | x0x1 [W] (int)
  __shared__ float red[4 * 8];
  float* s_src = sm;
  float* s_wy = s_src + h * w;
  float* s_wx = s_wy + h;
  float* s_fx = s_wx + w;
  int* s_x01 = reinterpret_cast<int*>(s_fx + W);
  const int g = blockIdx.x / T, t = blockIdx.x % T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

  for (int i = tid; i < h + w; i += blockDim.x) s_wy[i] = 0.f;  // wy and wx are contiguous
  __syncthreads();
  for (int y = tid; y < H; y += blockDim.x) {
    int y0, y1; float fy;
    src_coord(y, h, H, y0, y1, fy);
    atomicAdd(&s_wy[y0], 1.f - fy);
    atomicAdd(&s_wy[y1], fy);
  }
  for (int x = tid; x < W; x += blockDim.x) {
    int x0, x1; float fx;
    src_coord(x, w, W, x0, x1, fx);
    atomicAdd(&s_wx[x0], 1.f - fx);
    atomicAdd(&s_wx[x1], fx);
    s_fx[x] = fx;
    s_x01[x] = x0 | (x1 << 16);
  }

  for (int ch = 0; ch < nch; ++ch) {
    const float* src = masks + (((long long)g * nch + ch) * T + t) * (long long)(h * w);
    __syncthreads();  // previous channel done with s_src (and, first time, the weight tables are complete)
    float lmax = -INFINITY;
    for (int i = tid; i < h * w; i += blockDim.x) {
      const float v = src[i];
      s_src[i] = v;
      lmax = fmaxf(lmax, v);
    }
    if (ch == 0) {
      lmax = warp_max(lmax);
      if (lane == 0) red[warp] = lmax;
    }
    __syncthreads();
    if (ch == 0) {
      float m = red[0];
      for (int i = 1; i < nwarps; ++i) m = fmaxf(m, red[i]);
      const float ml2 = m * 1.4426950408889634f;
      float s = 0.f, sx = 0.f, sy = 0.f;
      for (int y = warp; y < H; y += nwarps) {
        int y0, y1; float fy;
        src_coord(y, h, H, y0, y1, fy);
        const float* r0 = s_src + y0 * w;
        const float* r1 = s_src + y1 * w;
        float rs = 0.f, rsx = 0.f;
        for (int x = lane; x < W; x += 32) {
          const int x01 = s_x01[x];
          const int x0 = x01 & 0xffff, x1 = x01 >> 16;
          const float fx = s_fx[x];
          const float top = fmaf(fx, r0[x1] - r0[x0], r0[x0]);
          const float bot = fmaf(fx, r1[x1] - r1[x0], r1[x0]);
          const float v = fmaf(fy, bot - top, top);
          const float e = ex2(fmaf(v, 1.4426950408889634f, -ml2));
          rs += e;
          rsx = fmaf(e, (float)x + 0.5f, rsx);
        }
        s += rs;
        sx += rsx;
        sy = fmaf(rs, (float)y + 0.5f, sy);
      }
      s = warp_sum(s); sx = warp_sum(sx); sy = warp_sum(sy);
      __syncthreads();  // everybody has read red[] (the maximum)
      if (lane == 0) { red[warp] = s; red[8 + warp] = sx; red[16 + warp] = sy; }
      __syncthreads();
      if (tid == 0) {
        float ts = 0.f, tx = 0.f, ty = 0.f;
        for (int i = 0; i < nwarps; ++i) { ts += red[i]; tx += red[8 + i]; ty += red[16 + i]; }
        traj[((long long)g * 2 + 0) * T + t] = tx / ts;
        traj[((long long)g * 2 + 1) * T + t] = ty / ts;
      }
    } else {
      float a = 0.f;
      for (int i = tid; i < h * w; i += blockDim.x) {
        const int sy_ = i / w, sx_ = i - sy_ * w;
        a = fmaf(s_src[i], s_wy[sy_] * s_wx[sx_], a);
      }
      a = warp_sum(a);
      __syncthreads();
      if (lane == 0) red[warp] = a;
      __syncthreads();
      if (tid == 0) {
        float tot = 0.f;
        for (int i = 0; i < nwarps; ++i) tot += red[i];
        const float mean = tot / (float)(H * W);
        if (ch == 1 && has_vis) vis[(long long)g * T + t] = mean;
        else if (has_depth) depth[(long long)g * T + t] = expf(mean);
      }
    }
  }
}

}  // namespace l4p

using namespace l4p;

extern "C" int l4p_token_attention(const float* q, const void* k16, const void* v16, float* out, int G, int nq, int Nk,
                                   int H, int d, int64_t kv_group_rows, float scale, int bf16, void* stream) {
  L4P_REQUIRE(q && k16 && v16 && out, L4P_ERR_ARG, "l4p_token_attention: null pointer");
  L4P_REQUIRE(G > 0 && nq > 0 && nq <= kTokMaxQ && Nk > 0 && H > 0 && d % 8 == 0 && d <= 64 * kTokMaxPairs, L4P_ERR_SHAPE,
              "l4p_token_attention: nq=%d (<=8) Nk=%d d=%d (multiple of 8, <=192)", nq, Nk, d);
  const size_t smem = sizeof(float) * ((size_t)nq * Nk + (size_t)nq * d + 8 * (size_t)nq * d);
  L4P_REQUIRE(smem <= 200 * 1024, L4P_ERR_SHAPE, "l4p_token_attention: Nk=%d too large", Nk);
  if (d == 88 && nq == 6 && Nk <= 2048 && Nk >= 256) {
    const size_t sm2 = sizeof(float) * ((size_t)6 * Nk + 6 * 88 + 16 * 6 * 88);
    auto kf = bf16 ? token_attention_fast_kernel<true, 88, 6> : token_attention_fast_kernel<false, 88, 6>;
    L4P_CHECK_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kf<<<dim3(H, G), 256, sm2, (cudaStream_t)stream>>>(q, (const uint16_t*)k16, (const uint16_t*)v16, out, Nk, H * d, H * d,
                                                        kv_group_rows, scale);
    L4P_CHECK_CUDA(cudaGetLastError());
    return L4P_OK;
  }
  auto kfn = bf16 ? token_attention_kernel<true> : token_attention_kernel<false>;
  L4P_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  kfn<<<dim3(H, G), 256, smem, (cudaStream_t)stream>>>(q, (const uint16_t*)k16, (const uint16_t*)v16, out, nq, Nk, d, H * d,
                                                       H * d, kv_group_rows, scale);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}


namespace l4p {
// ----------------------------------------------------------------------------------------------------------------------
// Streaming image -> token attention (round 2): persistent CTAs, 32-row tiles of the 16-bit query stream moved by bulk TMA
// copies through a 4-deep shared-memory ring (three tiles = 135 KB in flight per SM; the round-1 kernel above has one
// synchronous tile: 467 us = 0.24 of the HBM roofline, long-scoreboard bound), computed in place, written back with bulk
// stores. A first version kept the FMA formulation (thread = (row, head), K / V as fp32 in shared memory): 356 us = 0.32 -
// not HBM- but LDS-bound (see the tensor-core kernel below, which replaced it: 184 us = 0.62).
// ----------------------------------------------------------------------------------------------------------------------
constexpr int kIaRows = 32, kIaThreads = 256, kIaD = 88, kIaH = 8;
constexpr int kIaRowBytes = kIaH * kIaD * 2;              // 1408

L4P_DEVICE void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
L4P_DEVICE void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
L4P_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
L4P_DEVICE void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
L4P_DEVICE void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace l4p


namespace l4p {
// ----------------------------------------------------------------------------------------------------------------------
// Image -> token attention, tensor-core formulation (round 2, default). The limiter of the FMA kernels above is not HBM but
// the shared-memory -> register path: every (row, head) thread re-reads all of K and V (4.2 KB) for 1056 FMAs, and an LDS.128
// returns 512 bytes per warp whether or not the lanes broadcast (measured: 356 us = 0.32 of the HBM roofline even with three
// bulk-TMA tiles in flight). Here K^T (pre-scaled) and V of one head live in REGISTERS as mma.sync B fragments for the whole
// kernel (12 + 11 registers), the query tile is the A operand fetched with ldmatrix (each byte of the tile is read once), the
// probabilities stay in the accumulator layout, which is already the A layout of the P V product (the FA2 trick), and the
// result overwrites the query slice in place. Warp w = head w; a 32-row tile = two 16-row MMA units per warp.
//   scores  S[16 x 8 keys]  = Q[16 x 96] K^T[96 x 8]    6 x mma.m16n8k16   (dims 88..95 of a head's slice are the next head's
//                                                                          first channels: multiplied by zero rows of K^T)
//   output  O[16 x 88]      = P[16 x 8]  V[8 x 88]      11 x mma.m16n8k8
// Tiles move as 32 per-row bulk copies (1408 B each) into rows of pitch 1424 B, so that the eight 16-byte row segments of an
// ldmatrix phase fall into eight different bank groups; four tiles in the ring. K, V and P are rounded to the operand type
// (like the tcgen05 variant of round 1 and like every other contraction of the path); accumulation is fp32.
// ----------------------------------------------------------------------------------------------------------------------
constexpr int kImPitch = kIaRowBytes + 16;                 // 1424
constexpr int kImTileBytes = kIaRows * kImPitch;           // 45568
constexpr int kImStages = 4;

L4P_DEVICE void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool BF16>
L4P_DEVICE void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <bool BF16>
L4P_DEVICE void mma_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  if constexpr (BF16)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
  else
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}

template <bool BF16>
__global__ void __launch_bounds__(kIaThreads, 1)
image_attention_mma_kernel(const uint16_t* __restrict__ q16, const float* __restrict__ kf, const float* __restrict__ vf,
                           uint16_t* __restrict__ out16, int Np, int nk, float scale, long long n_tiles) {
  extern __shared__ __align__(128) uint8_t im_smem[];
  __shared__ __align__(8) uint64_t bar_full[kImStages], bar_done[kImStages];
  const uint32_t ring = smem_u32(im_smem);
  constexpr int ld = kIaH * kIaD;   // 704 elements per row
  const long long per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * per;
  const long long t_end = t_begin + per < n_tiles ? t_begin + per : n_tiles;
  const int tiles_per_group = Np / kIaRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kImStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_done[s]), kIaThreads / 32);   // one arrival per warp: its head's slice of the tile is final
    }
    fence_mbar_init();
  }
  // the 16 pad bytes of every ring row are read (times zero) by the last k-step of head 7: never leave NaN patterns there
  for (int i = threadIdx.x; i < kImStages * kIaRows; i += kIaThreads)
    *reinterpret_cast<uint4*>(im_smem + (size_t)i * kImPitch + kIaRowBytes) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  __syncthreads();
  if (t_begin >= t_end) return;
  const int n_local = (int)(t_end - t_begin);
  auto load_tile = [&](int i) {   // warp 0, all lanes: lane = row
    const int s = i % kImStages;
    const uint32_t fb = smem_u32(&bar_full[s]);
    if (lane == 0) mbar_expect_tx(fb, kIaRows * kIaRowBytes);
    __syncwarp();
    bulk_load(ring + (uint32_t)s * kImTileBytes + (uint32_t)lane * kImPitch,
              q16 + ((t_begin + i) * (long long)kIaRows + lane) * ld, kIaRowBytes, fb);
  };
  if (warp == 0)
    for (int i = 0; i < kImStages - 1 && i < n_local; ++i) load_tile(i);

  const int h = warp;               // head of this warp
  const int g4 = lane >> 2, c4 = lane & 3;
  uint32_t kfrag[6][2], vfrag[11];
  int cur_g = -1;
  for (int i = 0; i < n_local; ++i) {
    const long long tile = t_begin + i;
    const int g = (int)(tile / tiles_per_group);
    if (g != cur_g) {   // this group's K^T / V B-fragments (rare: a group spans Np / 32 consecutive tiles)
      const float* kg = kf + (long long)g * nk * ld + h * kIaD;
      const float* vg = vf + (long long)g * nk * ld + h * kIaD;
#pragma unroll
      for (int ks = 0; ks < 6; ++ks)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int d0 = ks * 16 + half * 8 + 2 * c4;   // B[k = dim][n = key g4]
          float a = 0.f, b = 0.f;
          if (g4 < nk && d0 < kIaD) { a = kg[g4 * ld + d0] * scale; b = kg[g4 * ld + d0 + 1] * scale; }
          kfrag[ks][half] = pack2<BF16>(a, b);
        }
#pragma unroll
      for (int nt = 0; nt < 11; ++nt) {                  // B[k = key 2 c4 (+1)][n = dim 8 nt + g4]
        const int j0 = 2 * c4, dd = nt * 8 + g4;
        const float a = j0 < nk ? vg[j0 * ld + dd] : 0.f, b = j0 + 1 < nk ? vg[(j0 + 1) * ld + dd] : 0.f;
        vfrag[nt] = pack2<BF16>(a, b);
      }
      cur_g = g;
    }
    const int s = i % kImStages;
    mbar_wait(smem_u32(&bar_full[s]), (uint32_t)(i / kImStages) & 1u);
    const uint32_t tbase = ring + (uint32_t)s * kImTileBytes + (uint32_t)h * (kIaD * 2);
#pragma unroll
    for (int u = 0; u < 2; ++u) {                         // 16-row units
      const uint32_t ubase = tbase + (uint32_t)(u * 16) * kImPitch;
      // ldmatrix.x4 row addresses: lanes 0-7 rows 0-7, 8-15 rows 8-15 (dims +0), 16-23 rows 0-7, 24-31 rows 8-15 (dims +8)
      const uint32_t lrow = ubase + (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * kImPitch + (uint32_t)((lane >> 4) & 1) * 16;
      float sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(lrow + ks * 32, a0, a1, a2, a3);
        mma_16816<BF16>(sc, a0, a1, a2, a3, kfrag[ks][0], kfrag[ks][1]);
      }
      // softmax over the nk valid keys: this thread holds keys 2 c4, 2 c4 + 1 of rows g4 (sc[0..1]) and g4 + 8 (sc[2..3])
      const bool v0 = 2 * c4 < nk, v1 = 2 * c4 + 1 < nk;
      float m0 = fmaxf(v0 ? sc[0] : -INFINITY, v1 ? sc[1] : -INFINITY), m1 = fmaxf(v0 ? sc[2] : -INFINITY, v1 ? sc[3] : -INFINITY);
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      const float p00 = v0 ? __expf(sc[0] - m0) : 0.f, p01 = v1 ? __expf(sc[1] - m0) : 0.f;
      const float p10 = v0 ? __expf(sc[2] - m1) : 0.f, p11 = v1 ? __expf(sc[3] - m1) : 0.f;
      float l0 = p00 + p01, l1 = p10 + p11;
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float i0 = 1.f / l0, i1 = 1.f / l1;
      const uint32_t pa0 = pack2<BF16>(p00 * i0, p01 * i0), pa1 = pack2<BF16>(p10 * i1, p11 * i1);
      // all lanes of the warp have fetched their A fragments (ldmatrix is warp-synchronous): the slice may be overwritten
      const uint32_t orow0 = ubase + (uint32_t)g4 * kImPitch + (uint32_t)c4 * 4, orow1 = orow0 + 8u * kImPitch;
#pragma unroll
      for (int nt = 0; nt < 11; ++nt) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        mma_1688<BF16>(o, pa0, pa1, vfrag[nt]);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(orow0 + nt * 16), "r"(pack2<BF16>(o[0], o[1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(orow1 + nt * 16), "r"(pack2<BF16>(o[2], o[3])) : "memory");
      }
    }
    // no CTA-wide barrier per tile: every warp publishes its slice (generic-proxy writes -> async proxy) and moves on to
    // the next tile of the ring; only warp 0, which owns the bulk copies, waits until all eight heads of the tile are final
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_done[s]));
    if (warp == 0) {
      mbar_wait(smem_u32(&bar_done[s]), (uint32_t)(i / kImStages) & 1u);
      bulk_store(out16 + (tile * (long long)kIaRows + lane) * ld, ring + (uint32_t)s * kImTileBytes + (uint32_t)lane * kImPitch,
                 kIaRowBytes);
      bulk_commit();
      const int nxt = i + kImStages - 1;   // reuses the slot of tile i - 1: this lane's store of ITS row of that tile must have
      if (nxt < n_local) {                 // finished reading shared memory before this lane's load overwrites the row
        bulk_wait_read<1>();
        load_tile(nxt);
      }
    }
  }
  if (warp == 0) bulk_wait_all<0>();
}
}  // namespace l4p

extern "C" int l4p_image_attention(const void* q16, const float* k, const float* v, void* out16, int G, int Np, int nk,
                                   int H, int d, float scale, int bf16, void* stream) {
  L4P_REQUIRE(q16 && k && v && out16, L4P_ERR_ARG, "l4p_image_attention: null pointer");
  L4P_REQUIRE(d == 88, L4P_ERR_SHAPE, "l4p_image_attention: head_dim=%d (this build: 88)", d);
  L4P_REQUIRE(G > 0 && nk > 0 && nk <= kTokMaxQ && H > 0 && H <= 8, L4P_ERR_SHAPE, "l4p_image_attention: nk=%d H=%d (<= 8)", nk, H);
  // L4P_IMGATT_STREAM=0 selects the round-1 kernel (A/B runs; it also serves shapes the streaming kernel does not take)
  static int stream_mode = -1;
  if (stream_mode < 0) {
    const char* e = getenv("L4P_IMGATT_STREAM");
    stream_mode = e ? atoi(e) : 1;
  }
  if (stream_mode != 0 && H == kIaH && Np % kIaRows == 0 && nk <= 8) {
    const long long n_tiles = (long long)G * Np / kIaRows;
    const size_t smem_m = (size_t)kImStages * kImTileBytes;
    typedef void (*MFn)(const uint16_t*, const float*, const float*, uint16_t*, int, int, float, long long);
    MFn mfn = bf16 ? image_attention_mma_kernel<true> : image_attention_mma_kernel<false>;
    static bool attr_m[2] = {false, false};
    if (!attr_m[bf16 ? 1 : 0]) {
      L4P_CHECK_CUDA(cudaFuncSetAttribute(mfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m));
      attr_m[bf16 ? 1 : 0] = true;
    }
    const long long sms = host_num_sms();
    const unsigned grid_m = (unsigned)(n_tiles < sms ? n_tiles : sms);
    mfn<<<grid_m, kIaThreads, smem_m, (cudaStream_t)stream>>>((const uint16_t*)q16, k, v, (uint16_t*)out16, Np, nk, scale, n_tiles);
    L4P_CHECK_CUDA(cudaGetLastError());
    return L4P_OK;
  }
  const int rows_per_block = 128;
  L4P_REQUIRE(Np % rows_per_block == 0, L4P_ERR_SHAPE, "l4p_image_attention: Np=%d must be a multiple of %d", Np, rows_per_block);
  L4P_REQUIRE(32 * H <= 256 && (H * d) % 8 == 0, L4P_ERR_SHAPE, "l4p_image_attention: H=%d", H);
  const size_t smem = sizeof(float) * 2 * (size_t)nk * H * d + 32 * (size_t)(H * d + 8) * 2;
  const unsigned grid = (unsigned)(((long long)G * Np) / rows_per_block);
  typedef void (*KFn)(const uint16_t*, const float*, const float*, uint16_t*, int, int, int, float, int);
  KFn kfn = bf16 ? image_attention_kernel<true, 88> : image_attention_kernel<false, 88>;
  L4P_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  kfn<<<grid, 32 * H, smem, (cudaStream_t)stream>>>((const uint16_t*)q16, k, v, (uint16_t*)out16, Np, nk, H, scale, rows_per_block);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}

extern "C" int l4p_layernorm16(const void* x16, const float* gamma, const float* beta, void* y16, int64_t rows, int cols,
                               float eps, int gelu, int bf16, void* stream) {
  L4P_REQUIRE(x16 && gamma && beta && y16, L4P_ERR_ARG, "l4p_layernorm16: null pointer");
  L4P_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0 && cols <= 2048, L4P_ERR_SHAPE, "l4p_layernorm16: cols=%d", cols);
  if (rows == 0) return L4P_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const uint16_t* x = (const uint16_t*)x16;
  uint16_t* y = (uint16_t*)y16;
  if (cols <= 8 * 8 * 6) {  // the mask decoder's 352-channel LayerNorm3d: 6 chunks per lane, 3 blocks per SM
    const unsigned grid = (unsigned)((rows + 31) / 32);
    if (bf16) layernorm16_kernel<true, 8, 6><<<grid, 256, 0, st>>>(x, gamma, beta, y, rows, cols, eps, gelu);
    else layernorm16_kernel<false, 8, 6><<<grid, 256, 0, st>>>(x, gamma, beta, y, rows, cols, eps, gelu);
  } else if (cols <= 8 * 8 * kLn16Iters) {  // short rows: 8 lanes per row, 4 rows per warp
    const unsigned grid = (unsigned)((rows + 31) / 32);
    if (bf16) layernorm16_kernel<true, 8, kLn16Iters><<<grid, 256, 0, st>>>(x, gamma, beta, y, rows, cols, eps, gelu);
    else layernorm16_kernel<false, 8, kLn16Iters><<<grid, 256, 0, st>>>(x, gamma, beta, y, rows, cols, eps, gelu);
  } else if (cols <= 32 * 8 * 6 && rows >= 4096) {  // the 1408-channel token stream: 4 rows per warp kept as raw words
    const unsigned grid = (unsigned)((rows + 15) / 16);
    if (bf16) layernorm16_rows_kernel<true, 6, 4><<<grid, 128, 0, st>>>(x, gamma, beta, y, rows, cols, eps, gelu);
    else layernorm16_rows_kernel<false, 6, 4><<<grid, 128, 0, st>>>(x, gamma, beta, y, rows, cols, eps, gelu);
  } else {
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (bf16) layernorm16_kernel<true, 32, kLn16Iters><<<grid, 256, 0, st>>>(x, gamma, beta, y, rows, cols, eps, gelu);
    else layernorm16_kernel<false, 32, kLn16Iters><<<grid, 256, 0, st>>>(x, gamma, beta, y, rows, cols, eps, gelu);
  }
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}

extern "C" int l4p_track_readout(const float* masks, float* traj, float* vis, float* depth, int G, int nch, int T, int h,
                                 int w, int H, int W, void* stream) {
  L4P_REQUIRE(masks && traj, L4P_ERR_ARG, "l4p_track_readout: null pointer");
  L4P_REQUIRE(G > 0 && nch >= 1 && nch <= 3 && T > 0 && h > 0 && w > 0 && H > 0 && W > 0 && ((size_t)h * w + h + w + 2 * (size_t)W) * 4 <= 160 * 1024,
              L4P_ERR_SHAPE, "l4p_track_readout: bad shape");
  L4P_REQUIRE((nch < 2 || vis) && (nch < 3 || depth), L4P_ERR_ARG, "l4p_track_readout: missing output");
  L4P_CHECK_CUDA(cudaFuncSetAttribute(track_readout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  L4P_REQUIRE(h < 65536 && w < 65536, L4P_ERR_SHAPE, "l4p_track_readout: source too large");
  track_readout_kernel<<<G * T, 256, ((size_t)h * w + h + w + 2 * (size_t)W) * 4, (cudaStream_t)stream>>>(masks, traj, vis, depth, T, h, w, H, W,
                                                                                nch >= 2, nch >= 3, nch);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}
