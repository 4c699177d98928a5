// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3-D convolution for sm_100a.
//
//   D[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
//   warp 0      TMA producer   (A/B tiles -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      UMMA issuer    (tcgen05.mma kind::f16, fp32 accumulators in TMEM, 2 accumulator stages)
//   warp 2      TMEM allocator
//   warps 4..11 epilogue       (tcgen05.ld -> bias/activation/residual -> global stores); two warpgroups interleave
//               16-column chunks so that the global-latency-bound epilogue of one overlaps the other
//
// A-operand modes:
//   A_MATRIX  plain row-major [M,K] matrix, one 2-D TMA box (64 x 128) per k-block.
//   A_CONV3D  channels-last activations [B,T,H,W,C]; a 128-row tile is a (bT,bH,bW) voxel box and the k-loop
//             runs over (filter tap, 64-channel block): every tap is the same 5-D TMA box shifted by the tap
//             offset, the zero padding comes from TMA out-of-bounds fill. No im2col buffer exists.
//
// Reference call sites replaced: modeling_finetune.py:62-69,171-177,188 (Linear), dpt_block.py:29-90,
// 144-157,255-278,406-414 (Conv3d / ConvTranspose3d), sam/transformer.py:223-245.
#include "common.cuh"
#include "../../include/l4p_b200.h"

namespace l4p {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // 64 x 2 B = one 128 B swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kAccCols = 256;                     // TMEM columns per accumulator stage
constexpr int kEpiGroups = 2;                     // epilogue warpgroups: group g handles 16-column chunks c with c % kEpiGroups == g
constexpr int kGemmThreads = 128 + 128 * kEpiGroups;

struct GemmKParams {
  int M, N, num_kb, block_n, stages;
  int tiles_m, tiles_n;
  int a_mode;
  // conv geometry
  int cB, cT, cH, cW, cCin, kT, kH, kW, bT, bH, bW, ntT, ntH, ntW, cblocks;
  // epilogue
  const float* bias;
  int act;
  const float* res_f32;
  const uint16_t* res_16;
  const uint16_t* res2_16;
  long long ld_res;
  int res_row_mod;
  int store_mode;
  float* out_f32;
  uint16_t* out_16;
  uint16_t* out_16_relu;
  long long ld_out;
  uint16_t *q, *k, *vt;
  int heads, head_dim, head_dim_pad, tokens;
  int sT, sH, sW, ctCout;
  const float* w2;
  const float* b2;
  int c2, exp_out;
  long long rows_per_group;  // STORE_HYPER: w2 is indexed by row / rows_per_group
};

struct TileCoord {
  int m_blk, n_blk;
  int b, t0, h0, w0;  // conv mode
};

L4P_DEVICE TileCoord decode_block(const GemmKParams& p, int m_blk, int n_blk) {
  TileCoord c;
  c.n_blk = n_blk;
  c.m_blk = m_blk;
  c.b = c.t0 = c.h0 = c.w0 = 0;
  if (p.a_mode == L4P_A_CONV3D) {
    int r = c.m_blk;
    c.w0 = (r % p.ntW) * p.bW; r /= p.ntW;
    c.h0 = (r % p.ntH) * p.bH; r /= p.ntH;
    c.t0 = (r % p.ntT) * p.bT; r /= p.ntT;
    c.b = r;  // may be >= cB for the padding block of an odd tile count (2-CTA mode): TMA zero-fills, rows are masked
  }
  return c;
}
L4P_DEVICE TileCoord decode_tile(const GemmKParams& p, int tile) {
  return decode_block(p, tile / p.tiles_n, tile % p.tiles_n);
}

// One 128-row x block_n accumulator tile: TMEM -> registers -> bias / activation / residual -> global memory.
// Shared by the 1-CTA and the 2-CTA (cta_group::2) kernels; `release` hands the accumulator stage back to the MMA warp.
template <bool BF16, class Release>
L4P_DEVICE void epilogue_tile(const GemmKParams& p, const TileCoord& tc, const int q4, const int lane, const int egrp,
                              const uint32_t tfull_bar, const uint32_t tfull_phase, const uint32_t t_acc, float* s_head,
                              Release release) {
  const int r = q4 * 32 + lane;  // row inside the tile
  long long row;                 // logical output row
  bool row_ok;
  int cb_ = 0, ct_ = 0, ch_ = 0, cw_ = 0;
  if (p.a_mode == L4P_A_CONV3D) {
    const int wl = r % p.bW;
    const int hl = (r / p.bW) % p.bH;
    const int tl = r / (p.bW * p.bH);
    ct_ = tc.t0 + tl; ch_ = tc.h0 + hl; cw_ = tc.w0 + wl; cb_ = tc.b;
    row_ok = (ct_ < p.cT) && (ch_ < p.cH) && (cw_ < p.cW) && (cb_ < p.cB);
    row = (((long long)cb_ * p.cT + ct_) * p.cH + ch_) * p.cW + cw_;
  } else {
    row = (long long)tc.m_blk * kBlockM + r;
    row_ok = row < p.M;
  }
  const int n0 = tc.n_blk * p.block_n;

  mbar_wait(tfull_bar, tfull_phase);
  tc_fence_after();
  const uint32_t t_addr = t_acc + ((uint32_t)(q4 * 32) << 16);

  float head_acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) head_acc[c] = 0.f;

  const long long rrow = p.res_row_mod > 0 ? row % p.res_row_mod : row;
  const float* hyper_w = p.store_mode == L4P_STORE_HYPER ? p.w2 + (row / p.rows_per_group) * (long long)(p.c2 * p.ctCout) : nullptr;
  for (int c0 = egrp * 16; c0 < p.block_n; c0 += 16 * kEpiGroups) {
    uint32_t raw[16];
    __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the per-row store predicate
    tmem_ld16(t_addr + (uint32_t)c0, raw);
    tmem_ld_wait();
    const int col0 = n0 + c0;
    if (col0 >= p.N) continue;  // uniform across the CTA
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
    if (p.bias != nullptr) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 bv = *reinterpret_cast<const float4*>(p.bias + col0 + i);
        v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
      }
    }
    if (p.act == L4P_ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = gelu_erf_fast(v[i]);
    } else if (p.act == L4P_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (!row_ok) {
      // masked row: nothing to store (loads above are warp-collective, stores are per-thread)
    } else if (p.store_mode == L4P_STORE_ROWMAJOR) {
      if (p.res_f32 != nullptr) {
        const float* rp = p.res_f32 + rrow * p.ld_res + col0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 rv = *reinterpret_cast<const float4*>(rp + i);
          v[i] += rv.x; v[i + 1] += rv.y; v[i + 2] += rv.z; v[i + 3] += rv.w;
        }
      }
      if (p.res_16 != nullptr) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res_16 + row * p.ld_res + col0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 rv = rp[h];
          const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = unpack2<BF16>(w[i]);
            v[h * 8 + 2 * i] += f.x; v[h * 8 + 2 * i + 1] += f.y;
          }
        }
      }
      if (p.res2_16 != nullptr) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res2_16 + row * p.ld_res + col0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 rv = rp[h];
          const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = unpack2<BF16>(w[i]);
            v[h * 8 + 2 * i] += f.x; v[h * 8 + 2 * i + 1] += f.y;
          }
        }
      }
      if (p.out_f32 != nullptr) {
        float* op = p.out_f32 + row * p.ld_out + col0;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      if (p.out_16 != nullptr) {
        uint4* op = reinterpret_cast<uint4*>(p.out_16 + row * p.ld_out + col0);
        op[0] = make_uint4(pack2<BF16>(v[0], v[1]), pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]),
                           pack2<BF16>(v[6], v[7]));
        op[1] = make_uint4(pack2<BF16>(v[8], v[9]), pack2<BF16>(v[10], v[11]), pack2<BF16>(v[12], v[13]),
                           pack2<BF16>(v[14], v[15]));
      }
      if (p.out_16_relu != nullptr) {
        uint4* op = reinterpret_cast<uint4*>(p.out_16_relu + row * p.ld_out + col0);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        op[0] = make_uint4(pack2<BF16>(v[0], v[1]), pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]),
                           pack2<BF16>(v[6], v[7]));
        op[1] = make_uint4(pack2<BF16>(v[8], v[9]), pack2<BF16>(v[10], v[11]), pack2<BF16>(v[12], v[13]),
                           pack2<BF16>(v[14], v[15]));
      }
    } else if (p.store_mode == L4P_STORE_QKV) {
      const int D = p.heads * p.head_dim;
      const long long bidx = row / p.tokens;
      const int tok = (int)(row - bidx * p.tokens);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = col0 + g * 8;
        const int s = col / D;
        const int rem = col - s * D;
        const int h = rem / p.head_dim;
        const int e = rem - h * p.head_dim;
        const long long bh = bidx * p.heads + h;
        const float* vv = v + g * 8;
        if (s < 2) {
          uint16_t* dst = (s == 0 ? p.q : p.k) + (bh * p.tokens + tok) * p.head_dim_pad + e;
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack2<BF16>(vv[0], vv[1]), pack2<BF16>(vv[2], vv[3]),
                                                      pack2<BF16>(vv[4], vv[5]), pack2<BF16>(vv[6], vv[7]));
        } else {
          uint16_t* dst = p.vt + (bh * p.head_dim_pad + e) * (long long)p.tokens + tok;
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[(long long)i * p.tokens] = pack1<BF16>(vv[i]);
        }
      }
    } else if (p.store_mode == L4P_STORE_CONVT) {
      // row = input voxel (b,t,h,w) of the [cB,cT,cH,cW] grid; col = ((kt*sH+kh)*sW+kw)*Cout + co
      long long rr = row;
      const int w_ = (int)(rr % p.cW); rr /= p.cW;
      const int h_ = (int)(rr % p.cH); rr /= p.cH;
      const int t_ = (int)(rr % p.cT); rr /= p.cT;
      const long long b_ = rr;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = col0 + g * 8;
        const int tapi = col / p.ctCout;
        const int co = col - tapi * p.ctCout;
        const int kw = tapi % p.sW;
        const int kh = (tapi / p.sW) % p.sH;
        const int kt = tapi / (p.sW * p.sH);
        const long long vox = ((b_ * (p.cT * p.sT) + (t_ * p.sT + kt)) * (p.cH * p.sH) + (h_ * p.sH + kh)) *
                                  (long long)(p.cW * p.sW) + (w_ * p.sW + kw);
        const float* vv = v + g * 8;
        *reinterpret_cast<uint4*>(p.out_16 + vox * p.ctCout + co) =
            make_uint4(pack2<BF16>(vv[0], vv[1]), pack2<BF16>(vv[2], vv[3]), pack2<BF16>(vv[4], vv[5]),
                       pack2<BF16>(vv[6], vv[7]));
      }
    } else if (p.store_mode == L4P_STORE_HYPER) {
      // v = act(acc + bias) of one ConvT tap (this N tile); dot with the per-query hyper-network vectors
      const float* wg = hyper_w + c0;
      float4 wv[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < p.c2)
#pragma unroll
          for (int i = 0; i < 4; ++i) wv[c][i] = *reinterpret_cast<const float4*>(wg + c * p.ctCout + 4 * i);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < p.c2) {
          float a = head_acc[c];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            a = fmaf(v[4 * i], wv[c][i].x, a); a = fmaf(v[4 * i + 1], wv[c][i].y, a);
            a = fmaf(v[4 * i + 2], wv[c][i].z, a); a = fmaf(v[4 * i + 3], wv[c][i].w, a);
          }
          head_acc[c] = a;
        }
      }
    } else {  // L4P_STORE_HEAD1X1: v already bias+ReLU'd; accumulate the tiny second conv
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < p.c2) {
          const float* wr = p.w2 + (long long)c * p.N + col0;
          float a = head_acc[c];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(wr + i);
            a = fmaf(v[i], wv.x, a); a = fmaf(v[i + 1], wv.y, a);
            a = fmaf(v[i + 2], wv.z, a); a = fmaf(v[i + 3], wv.w, a);
          }
          head_acc[c] = a;
        }
      }
    }
  }
  // accumulator stage drained -> hand TMEM back to the MMA warp
  tc_fence_before();
  release();

  if (p.store_mode == L4P_STORE_HYPER || p.store_mode == L4P_STORE_HEAD1X1) {
    // the per-row dot products were accumulated per warpgroup over its chunks: reduce them in group 0
    if (egrp != 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c) s_head[(egrp - 1) * 128 * 8 + r * 8 + c] = head_acc[c];
    }
    named_bar_sync(1, 128 * kEpiGroups);
    if (egrp == 0) {
#pragma unroll
      for (int g2 = 1; g2 < kEpiGroups; ++g2)
#pragma unroll
        for (int c = 0; c < 8; ++c) head_acc[c] += s_head[(g2 - 1) * 128 * 8 + r * 8 + c];
    }
    named_bar_sync(1, 128 * kEpiGroups);  // s_head may be overwritten by the next tile
  }

  if (p.store_mode == L4P_STORE_HYPER && row_ok && egrp == 0) {
    // row = input voxel (g,t,h,w) of the [cB,cT,cH,cW] grid; this N tile = tap (kt,kh,kw)
    long long rr = row;
    const int w_ = (int)(rr % p.cW); rr /= p.cW;
    const int h_ = (int)(rr % p.cH); rr /= p.cH;
    const int t_ = (int)(rr % p.cT); rr /= p.cT;
    const long long g_ = rr;
    const int tapi = tc.n_blk;
    const int kw = tapi % p.sW, kh = (tapi / p.sW) % p.sH, kt = tapi / (p.sW * p.sH);
    const long long oT = (long long)p.cT * p.sT, oH = (long long)p.cH * p.sH, oW = (long long)p.cW * p.sW;
    const long long vox = ((long long)(t_ * p.sT + kt) * oH + (h_ * p.sH + kh)) * oW + (w_ * p.sW + kw);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < p.c2) p.out_f32[((g_ * p.c2 + c) * oT * oH * oW) + vox] = head_acc[c];
  }
  if (p.store_mode == L4P_STORE_HEAD1X1 && row_ok && egrp == 0) {
    const long long plane = (long long)p.cT * p.cH * p.cW;
    const long long vox = ((long long)ct_ * p.cH + ch_) * p.cW + cw_;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < p.c2) {
        float o = head_acc[c] + p.b2[c];
        if (p.exp_out) o = expf(o);
        p.out_f32[((long long)cb_ * p.c2 + c) * plane + vox] = o;
      }
    }
  }
}

template <bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_head[(kEpiGroups - 1) * 128 * 8];  // cross-warpgroup reduction of the fused head dot products

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.block_n * 128u;
  const uint32_t stage_bytes = kABytes + b_bytes;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 128 * kEpiGroups);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int n0 = tc.n_blk * p.block_n;
        // filter-tap walk (cb fastest, then dw, dh, dt) kept as counters: no divisions in the single-thread hot loop
        int cb = 0, dw = -(p.kW / 2), dh = -(p.kH / 2), dt = -(p.kT / 2);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + kABytes;
          mbar_expect_tx(full, stage_bytes);
          if (p.a_mode == L4P_A_MATRIX) {
            tma_load_2d(sa, &tmA, full, kb * kBlockK, tc.m_blk * kBlockM);
          } else {
            tma_load_5d(sa, &tmA, full, cb * kBlockK, tc.w0 + dw, tc.h0 + dh, tc.t0 + dt, tc.b);
            if (++cb == p.cblocks) {
              cb = 0;
              if (++dw > p.kW / 2) {
                dw = -(p.kW / 2);
                if (++dh > p.kH / 2) { dh = -(p.kH / 2); ++dt; }
              }
            }
          }
          tma_load_2d(sb, &tmB, full, kb * kBlockK, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    // whole warp walks the pipeline (warp-uniform control flow), one elected lane issues
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_f16(BF16, kBlockM, (uint32_t)p.block_n);
    constexpr uint32_t hi128 = umma_desc_hi(128, 2);
    const uint32_t a_lo0 = umma_desc_lo(smem_base), b_lo0 = umma_desc_lo(smem_base + kABytes);
    const uint32_t stage_step = stage_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        if (leader) {
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_step, b_lo = b_lo0 + (uint32_t)stage * stage_step;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle row: +2 in the (addr >> 4) field
            umma_ss(d_tmem, umma_desc_make(a_lo + 2 * k, hi128), umma_desc_make(b_lo + 2 * k, hi128), idesc,
                    (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[stage]));
          if (kb == p.num_kb - 1) umma_commit(smem_u32(&bar_tfull[acc]));
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q4 = warp & 3;  // TMEM lane quarter owned by this warp
    const int egrp = (warp - 4) >> 2;  // epilogue warpgroup: interleaved 16-column chunks
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const uint32_t tempty = smem_u32(&bar_tempty[acc]);
      epilogue_tile<BF16>(p, tc, q4, lane, egrp, smem_u32(&bar_tfull[acc]), acc_phase, tmem_base + (uint32_t)acc * kAccCols,
                          s_head, [&]() { mbar_arrive(tempty); });
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// 2-CTA variant: a CTA pair (cluster of 2, same TPC) owns a 256-row x block_n tile. Each CTA stages its own 128 rows
// of A and HALF of the B tile per k-block (32 KiB instead of 48 KiB per 128 output rows: 128-row tiles are
// L2->SM-bandwidth-bound), the leader issues tcgen05.mma.cta_group::2 (M=256) which reads both CTAs' shared memory and
// writes each CTA's half of the accumulator into that CTA's own TMEM; both epilogues run independently.
//   full[s]    lives in the leader: 1 arrival (leader's expect_tx of BOTH CTAs' bytes) + the bytes of all four TMA loads
//   empty[s]   one per CTA, released by a multicast tcgen05.commit
//   tfull[a]   one per CTA (multicast commit);   tempty[a] in the leader: all epilogue threads of both CTAs arrive
// ------------------------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_head[(kEpiGroups - 1) * 128 * 8];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t half_n = (uint32_t)p.block_n / 2;
  const uint32_t b_bytes = half_n * 128u;
  const uint32_t stage_bytes = kABytes + b_bytes;
  const int tiles_m2 = (p.tiles_m + 1) / 2;
  const int num_tiles = tiles_m2 * p.tiles_n;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 2 * 128 * kEpiGroups);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc2(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast / peer TMA signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n_blk = tile % p.tiles_n;
        const TileCoord tc = decode_block(p, (tile / p.tiles_n) * 2 + (int)rank, n_blk);
        const int n0 = n_blk * p.block_n + (int)(rank * half_n);
        int cb = 0, dw = -(p.kW / 2), dh = -(p.kH / 2), dt = -(p.kT / 2);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full_leader = mapa_shared(smem_u32(&bar_full[stage]), 0);
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + kABytes;
          if (is_leader) mbar_expect_tx(smem_u32(&bar_full[stage]), 2 * stage_bytes);
          if (p.a_mode == L4P_A_MATRIX) {
            tma2_load_2d(sa, &tmA, full_leader, kb * kBlockK, tc.m_blk * kBlockM);
          } else {
            tma2_load_5d(sa, &tmA, full_leader, cb * kBlockK, tc.w0 + dw, tc.h0 + dh, tc.t0 + dt, tc.b);
            if (++cb == p.cblocks) {
              cb = 0;
              if (++dw > p.kW / 2) {
                dw = -(p.kW / 2);
                if (++dh > p.kH / 2) { dh = -(p.kH / 2); ++dt; }
              }
            }
          }
          tma2_load_2d(sb, &tmB, full_leader, kb * kBlockK, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer (leader CTA only)
    if (is_leader) {
      const bool leader_lane = elect_one();
      const uint32_t idesc = umma_idesc_f16(BF16, 2 * kBlockM, (uint32_t)p.block_n);
      constexpr uint32_t hi128 = umma_desc_hi(128, 2);
      const uint32_t a_lo0 = umma_desc_lo(smem_base), b_lo0 = umma_desc_lo(smem_base + kABytes);
      const uint32_t stage_step = stage_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          if (leader_lane) {
            const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_step, b_lo = b_lo0 + (uint32_t)stage * stage_step;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma2_ss(d_tmem, umma_desc_make(a_lo + 2 * k, hi128), umma_desc_make(b_lo + 2 * k, hi128), idesc,
                       (kb | k) != 0 ? 1u : 0u);
            umma2_commit_mc(smem_u32(&bar_empty[stage]), 3);
            if (kb == p.num_kb - 1) umma2_commit_mc(smem_u32(&bar_tfull[acc]), 3);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int q4 = warp & 3;
    const int egrp = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const TileCoord tc = decode_block(p, (tile / p.tiles_n) * 2 + (int)rank, tile % p.tiles_n);
      const uint32_t tempty_leader = mapa_shared(smem_u32(&bar_tempty[acc]), 0);
      epilogue_tile<BF16>(p, tc, q4, lane, egrp, smem_u32(&bar_tfull[acc]), acc_phase, tmem_base + (uint32_t)acc * kAccCols,
                          s_head, [&]() { mbar_arrive_cluster(tempty_leader); });
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still signal / read this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

static int pick_block_n(long long N) {
  // prefer the widest tile that divides N (UMMA N <= 256, multiple of 16)
  const int cands[] = {256, 240, 224, 208, 192, 176, 160, 144, 128};
  for (int c : cands)
    if (N % c == 0) return c;
  if (N <= 256) return (int)((N + 15) / 16 * 16);
  return 256;  // ragged tail handled by TMA zero fill + column masking
}

}  // namespace l4p

using namespace l4p;

extern "C" int l4p_gemm(const l4p_gemm_desc* d, void* stream_) {
  L4P_REQUIRE(d != nullptr, L4P_ERR_ARG, "l4p_gemm: null descriptor");
  cudaStream_t stream = (cudaStream_t)stream_;
  L4P_REQUIRE(d->a && d->w, L4P_ERR_ARG, "l4p_gemm: null operand");
  L4P_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, L4P_ERR_SHAPE, "l4p_gemm: empty problem M=%lld N=%lld K=%lld",
              (long long)d->M, (long long)d->N, (long long)d->K);
  L4P_REQUIRE(d->N % 16 == 0, L4P_ERR_SHAPE, "l4p_gemm: N=%lld must be a multiple of 16", (long long)d->N);
  L4P_REQUIRE(d->ldw % 8 == 0 && d->ldw >= d->K, L4P_ERR_SHAPE, "l4p_gemm: ldw=%lld", (long long)d->ldw);

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)d->M;
  p.N = (int)d->N;
  p.block_n = d->block_n > 0 ? d->block_n : pick_block_n(d->N);
  if (d->block_n <= 0 && d->store_mode != L4P_STORE_HEAD1X1) {
    // few output tiles (low-resolution pyramid levels, token-side GEMMs): trade tile width for CTAs so that more
    // than a handful of SMs work on the (long) K loop
    const long long tm = d->a_mode == L4P_A_CONV3D
                             ? (long long)d->cB * ((d->cT + d->bT - 1) / d->bT) * ((d->cH + d->bH - 1) / d->bH) * ((d->cW + d->bW - 1) / d->bW)
                             : (d->M + kBlockM - 1) / kBlockM;
    while (p.block_n >= 64 && (p.block_n / 2) % 16 == 0 && tm * ((d->N + p.block_n - 1) / p.block_n) < 96)
      p.block_n /= 2;
  }
  L4P_REQUIRE(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, L4P_ERR_SHAPE, "l4p_gemm: block_n=%d",
              p.block_n);
  p.a_mode = d->a_mode;
  p.tiles_n = (int)((d->N + p.block_n - 1) / p.block_n);

  CUtensorMap tmA, tmB;
  int rc;
  if (d->a_mode == L4P_A_MATRIX) {
    L4P_REQUIRE(d->lda % 8 == 0 && d->lda >= d->K, L4P_ERR_SHAPE, "l4p_gemm: lda=%lld", (long long)d->lda);
    p.num_kb = (int)((d->K + kBlockK - 1) / kBlockK);
    p.tiles_m = (int)((d->M + kBlockM - 1) / kBlockM);
    const uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->M};
    const uint64_t strides[1] = {(uint64_t)d->lda * 2};
    const uint32_t box[2] = {kBlockK, kBlockM};
    rc = host_make_tmap_16b(&tmA, d->a, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  } else if (d->a_mode == L4P_A_CONV3D) {
    L4P_REQUIRE(d->cCin % kBlockK == 0, L4P_ERR_SHAPE, "l4p_gemm(conv): Cin=%d must be a multiple of 64", d->cCin);
    L4P_REQUIRE(d->bT * d->bH * d->bW == kBlockM, L4P_ERR_SHAPE, "l4p_gemm(conv): box %dx%dx%d != 128 voxels",
                d->bT, d->bH, d->bW);
    L4P_REQUIRE((d->kT & 1) && (d->kH & 1) && (d->kW & 1), L4P_ERR_SHAPE, "l4p_gemm(conv): even filter extent");
    L4P_REQUIRE(d->K == (int64_t)d->kT * d->kH * d->kW * d->cCin, L4P_ERR_SHAPE, "l4p_gemm(conv): K mismatch");
    L4P_REQUIRE(d->M == (int64_t)d->cB * d->cT * d->cH * d->cW, L4P_ERR_SHAPE, "l4p_gemm(conv): M mismatch");
    p.cB = d->cB; p.cT = d->cT; p.cH = d->cH; p.cW = d->cW; p.cCin = d->cCin;
    p.kT = d->kT; p.kH = d->kH; p.kW = d->kW;
    p.bT = d->bT; p.bH = d->bH; p.bW = d->bW;
    p.ntT = (d->cT + d->bT - 1) / d->bT;
    p.ntH = (d->cH + d->bH - 1) / d->bH;
    p.ntW = (d->cW + d->bW - 1) / d->bW;
    p.cblocks = d->cCin / kBlockK;
    p.num_kb = d->kT * d->kH * d->kW * p.cblocks;
    p.tiles_m = d->cB * p.ntT * p.ntH * p.ntW;
    const uint64_t C = (uint64_t)d->cCin;
    const uint64_t dims[5] = {C, (uint64_t)d->cW, (uint64_t)d->cH, (uint64_t)d->cT, (uint64_t)d->cB};
    const uint64_t strides[4] = {C * 2, C * 2 * d->cW, C * 2 * d->cW * d->cH, C * 2 * d->cW * d->cH * d->cT};
    const uint32_t box[5] = {kBlockK, (uint32_t)d->bW, (uint32_t)d->bH, (uint32_t)d->bT, 1};
    rc = host_make_tmap_16b(&tmA, d->a, 5, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  } else {
    return host_set_error(L4P_ERR_ARG, "l4p_gemm: a_mode=%d", d->a_mode);
  }
  {
    const uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->N};
    const uint64_t strides[1] = {(uint64_t)d->ldw * 2};
    const uint32_t box[2] = {kBlockK, (uint32_t)p.block_n};
    rc = host_make_tmap_16b(&tmB, d->w, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  }

  // epilogue validation
  p.bias = d->bias;
  p.act = d->act;
  p.store_mode = d->store_mode;
  p.res_f32 = d->res_f32;
  p.res_16 = (const uint16_t*)d->res_16;
  p.res2_16 = (const uint16_t*)d->res2_16;
  p.ld_res = d->ld_res;
  p.res_row_mod = d->res_row_mod;
  p.out_f32 = d->out_f32;
  p.out_16 = (uint16_t*)d->out_16;
  p.out_16_relu = (uint16_t*)d->out_16_relu;
  p.ld_out = d->ld_out;
  switch (d->store_mode) {
    case L4P_STORE_ROWMAJOR:
      L4P_REQUIRE(d->out_f32 || d->out_16 || d->out_16_relu, L4P_ERR_ARG, "l4p_gemm: no output");
      L4P_REQUIRE(d->ld_out % 8 == 0 && d->ld_out >= d->N, L4P_ERR_SHAPE, "l4p_gemm: ld_out=%lld", (long long)d->ld_out);
      if (d->res_f32 || d->res_16 || d->res2_16)
        L4P_REQUIRE(d->ld_res % 8 == 0 && d->ld_res >= d->N, L4P_ERR_SHAPE, "l4p_gemm: ld_res=%lld", (long long)d->ld_res);
      break;
    case L4P_STORE_QKV:
      L4P_REQUIRE(d->q && d->k && d->vt, L4P_ERR_ARG, "l4p_gemm(qkv): null q/k/vt");
      L4P_REQUIRE(d->head_dim % 8 == 0 && d->head_dim_pad % 8 == 0 && d->head_dim_pad >= d->head_dim, L4P_ERR_SHAPE,
                  "l4p_gemm(qkv): head_dim=%d pad=%d", d->head_dim, d->head_dim_pad);
      L4P_REQUIRE(d->N == 3ll * d->heads * d->head_dim, L4P_ERR_SHAPE, "l4p_gemm(qkv): N != 3*heads*head_dim");
      L4P_REQUIRE(d->tokens > 0 && d->M % d->tokens == 0 && d->tokens % 8 == 0, L4P_ERR_SHAPE, "l4p_gemm(qkv): tokens=%d",
                  d->tokens);
      p.q = (uint16_t*)d->q; p.k = (uint16_t*)d->k; p.vt = (uint16_t*)d->vt;
      p.heads = d->heads; p.head_dim = d->head_dim; p.head_dim_pad = d->head_dim_pad; p.tokens = d->tokens;
      break;
    case L4P_STORE_CONVT:
      L4P_REQUIRE(d->out_16, L4P_ERR_ARG, "l4p_gemm(convT): null out_16");
      L4P_REQUIRE(d->ctCout % 8 == 0 && d->N == (int64_t)d->sT * d->sH * d->sW * d->ctCout, L4P_ERR_SHAPE,
                  "l4p_gemm(convT): N != sT*sH*sW*Cout");
      L4P_REQUIRE(d->M == (int64_t)d->cB * d->cT * d->cH * d->cW, L4P_ERR_SHAPE, "l4p_gemm(convT): M mismatch");
      p.cB = d->cB; p.cT = d->cT; p.cH = d->cH; p.cW = d->cW;
      p.sT = d->sT; p.sH = d->sH; p.sW = d->sW; p.ctCout = d->ctCout;
      break;
    case L4P_STORE_HEAD1X1:
      L4P_REQUIRE(d->a_mode == L4P_A_CONV3D, L4P_ERR_ARG, "l4p_gemm(head1x1): conv mode only");
      L4P_REQUIRE(d->out_f32 && d->w2 && d->b2 && d->c2 >= 1 && d->c2 <= 8, L4P_ERR_ARG, "l4p_gemm(head1x1): args");
      L4P_REQUIRE(p.tiles_n == 1, L4P_ERR_SHAPE, "l4p_gemm(head1x1): N=%lld must fit one tile", (long long)d->N);
      p.w2 = d->w2; p.b2 = d->b2; p.c2 = d->c2; p.exp_out = d->exp_out;
      break;
    case L4P_STORE_HYPER:
      L4P_REQUIRE(d->out_f32 && d->w2 && d->c2 >= 1 && d->c2 <= 4, L4P_ERR_ARG, "l4p_gemm(hyper): args");
      L4P_REQUIRE(d->ctCout % 16 == 0 && d->N == (int64_t)d->sT * d->sH * d->sW * d->ctCout, L4P_ERR_SHAPE,
                  "l4p_gemm(hyper): N != sT*sH*sW*Cout");
      L4P_REQUIRE(p.block_n == d->ctCout, L4P_ERR_SHAPE, "l4p_gemm(hyper): block_n must equal Cout (one tap per tile)");
      L4P_REQUIRE(d->M == (int64_t)d->cB * d->cT * d->cH * d->cW && d->rows_per_group > 0, L4P_ERR_SHAPE,
                  "l4p_gemm(hyper): M mismatch");
      p.cB = d->cB; p.cT = d->cT; p.cH = d->cH; p.cW = d->cW;
      p.sT = d->sT; p.sH = d->sH; p.sW = d->sW; p.ctCout = d->ctCout;
      p.w2 = d->w2; p.c2 = d->c2; p.rows_per_group = d->rows_per_group;
      break;
    default:
      return host_set_error(L4P_ERR_ARG, "l4p_gemm: store_mode=%d", d->store_mode);
  }

  const uint32_t stage_bytes = kABytes + (uint32_t)p.block_n * 128u;
  int stages = (int)((216u * 1024u) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > p.num_kb) stages = p.num_kb < 2 ? 2 : p.num_kb;
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;

  const int num_tiles = p.tiles_m * p.tiles_n;
  int grid = host_num_sms();
  if (grid > num_tiles) grid = num_tiles;

  // CTA pairs (256-row tiles) when the problem is large enough to fill the machine with pair tiles
  const int pair_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;
  const int pairs = host_num_sms() / 2;
  int use_pair = d->cta_pair;  // 0 = auto, 1 = force, -1 = never
  if (use_pair == 0) use_pair = (pair_tiles >= pairs && p.block_n >= 64 && p.tiles_m >= 2) ? 1 : -1;
  if (use_pair == 1) {
    L4P_REQUIRE(p.block_n % 32 == 0 || p.block_n % 16 == 0, L4P_ERR_SHAPE, "l4p_gemm(pair): block_n=%d", p.block_n);
    // B tensor map with the half-tile box
    const uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->N};
    const uint64_t strides[1] = {(uint64_t)d->ldw * 2};
    const uint32_t box[2] = {kBlockK, (uint32_t)(p.block_n / 2)};
    rc = host_make_tmap_16b(&tmB, d->w, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
    const uint32_t sb2 = kABytes + (uint32_t)(p.block_n / 2) * 128u;
    int st2 = (int)((216u * 1024u) / sb2);
    if (st2 > kMaxStages) st2 = kMaxStages;
    if (st2 > p.num_kb) st2 = p.num_kb < 2 ? 2 : p.num_kb;
    if (st2 > kMaxStages) st2 = kMaxStages;
    p.stages = st2;
    const size_t smem2 = (size_t)st2 * sb2 + 1024;
    int g2 = pairs < pair_tiles ? pairs : pair_tiles;
    auto k2 = d->bf16 ? gemm2_kernel<true> : gemm2_kernel<false>;
    static bool attr2[2] = {false, false};
    if (!attr2[d->bf16 ? 1 : 0]) {
      L4P_CHECK_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));
      attr2[d->bf16 ? 1 : 0] = true;
    }
    k2<<<2 * g2, kGemmThreads, smem2, stream>>>(tmA, tmB, p);
    L4P_CHECK_CUDA(cudaGetLastError());
    return L4P_OK;
  }

  auto kfn = d->bf16 ? gemm_kernel<true> : gemm_kernel<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[d->bf16 ? 1 : 0]) {
    L4P_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));
    attr_set[d->bf16 ? 1 : 0] = true;
  }
  kfn<<<grid, kGemmThreads, smem, stream>>>(tmA, tmB, p);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}
