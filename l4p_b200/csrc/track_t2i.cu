// Token -> video-token attention of the track head's two-way transformer without the K / V projections of the video tokens.
//
// sam/transformer.py:223-245 computes, per track query g (G = 128 of them, each with its OWN 2048 video tokens x_n of 1408
// channels after the first two-way layer) and head h (8 heads of 88): K = W_k (x + pe) + b_k, V = W_v x + b_v for all 2048
// tokens, although only nt = 6 prompt tokens attend to them. Reference order: 2 x 2048 x 1408 x 704 MACs per query and site
// (2 x 0.52 TFLOP per launch at G = 128, 0.55 ms each on the tensor cores, three sites per window). Since the scores and
// the attention output are linear in K and V, the projections move to the 6-token side:
//     score[g,h,t,n] = q[g,t,h] . K[g,n,h]           = x[g,n] . (W_k[h]^T q[g,t,h])   +  q[g,t,h] . (W_k pe_n + b_k)[h]
//     out[g,t,h]     = sum_n p[g,h,t,n] V[g,n,h]     = W_v[h] (sum_n p[g,h,t,n] x[g,n]) + b_v[h]         (sum_n p = 1)
// i.e. per query a [48 x 1408] x [1408 x 2048] score GEMM (48 = 8 heads x 6 tokens; l4p_gemm with grouped weights), a row
// softmax, and a [48 x 2048] x [2048 x 1408] weighted token sum: 15 x fewer MACs, and the per-query token stream is read
// twice instead of being projected into two 704-wide copies that are written and read again.
//
//   l4p_head_expand         q fp32 [G*nt, H*hd] -> block-diagonal 16-bit operand [G*H*nt, H*hd] (row (h,t) keeps head h only), x scale
//   l4p_row_softmax16       fp32 scores [rows, n] -> 16-bit probabilities
//   l4p_token_weighted_sum  Y[g] = P[g] X[g]: mma.sync streaming kernel, X tiles through a TMA ring, fp32 accumulation
//   l4p_head_diag_gather    Z fp32 [G*H*nt, H*hd] -> out fp32 [G*nt, H*hd]: head h of row (h,t)
#include "common.cuh"
#include "../../include/l4p_b200.h"

namespace l4p {

template <bool BF16>
__global__ void __launch_bounds__(256)
head_expand_kernel(const float* __restrict__ q, uint16_t* __restrict__ out, long long total8, int nt, int heads, int hd, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int D = heads * hd, d8 = D >> 3, J = heads * nt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / d8;            // g * J + h * nt + t
    const int col = (int)(i - row * d8) * 8;
    const long long g = row / J;
    const int j = (int)(row - g * J);
    const int h = j / nt, t = j - h * nt;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (col / hd == h) {                     // hd is a multiple of 8: an 8-column group never straddles a head
      const float4* src = reinterpret_cast<const float4*>(q + (g * nt + t) * D + col);
      const float4 a = src[0], b = src[1];
      o = make_uint4(pack2<BF16>(a.x * scale, a.y * scale), pack2<BF16>(a.z * scale, a.w * scale),
                     pack2<BF16>(b.x * scale, b.y * scale), pack2<BF16>(b.z * scale, b.w * scale));
    }
    *reinterpret_cast<uint4*>(out + row * D + col) = o;
  }
}

__global__ void __launch_bounds__(256)
head_diag_gather_kernel(const float* __restrict__ z, float* __restrict__ out, long long total4, int nt, int heads, int hd) {
  pdl_launch_dependents();
  pdl_wait();
  const int D = heads * hd, d4 = D >> 2, J = heads * nt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / d4;            // g * nt + t
    const int col = (int)(i - row * d4) * 4;
    const long long g = row / nt;
    const int t = (int)(row - g * nt);
    const int h = col / hd;
    *reinterpret_cast<float4*>(out + row * D + col) = *reinterpret_cast<const float4*>(z + (g * J + h * nt + t) * D + col);
  }
}

// one warp per row, the row in registers (n <= 2048), exp2 on pre-scaled scores, probabilities rounded once when stored
constexpr int kRsVec = 16;  // float4 per lane
template <bool BF16>
__global__ void __launch_bounds__(256)
row_softmax16_kernel(const float* __restrict__ s, uint16_t* __restrict__ p, long long rows, int n) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const int nvec = n >> 2;
  const float4* sr = reinterpret_cast<const float4*>(s + row * n);
  float4 v[kRsVec];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kRsVec; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      v[i] = sr[j];
      mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
  }
  mx = warp_max(mx);
  const float L2E = 1.4426950408889634f;
  const float off = mx * L2E;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kRsVec; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      v[i].x = ex2(fmaf(v[i].x, L2E, -off)); v[i].y = ex2(fmaf(v[i].y, L2E, -off));
      v[i].z = ex2(fmaf(v[i].z, L2E, -off)); v[i].w = ex2(fmaf(v[i].w, L2E, -off));
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float inv = 1.0f / warp_sum(sum);
  uint2* pr = reinterpret_cast<uint2*>(p + row * n);
#pragma unroll
  for (int i = 0; i < kRsVec; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) pr[j] = make_uint2(pack2<BF16>(v[i].x * inv, v[i].y * inv), pack2<BF16>(v[i].z * inv, v[i].w * inv));
  }
}

// Video-token -> token attention, folded form: scores arrive TRANSPOSED, s fp32 [G*heads*nt, n] (row (g,h,t), column = video
// token), the softmax runs over the nt tokens of each head, and the probabilities leave as the A operand of the per-query
// K = heads*nt output GEMM: p16 [G*n, heads*nt]. One thread per (g, video token): its reads are coalesced across the
// warp (consecutive tokens of one score row), its heads*nt probabilities are contiguous in the output.
template <bool BF16, int NT>
__global__ void __launch_bounds__(256)
group_softmax_t16_kernel(const float* __restrict__ s, uint16_t* __restrict__ p, long long total, int n, int heads) {
  pdl_launch_dependents();
  pdl_wait();
  const int J = heads * NT;
  const float L2E = 1.4426950408889634f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / n;
    const int tok = (int)(i - g * n);
    const float* src = s + (g * J) * n + tok;
    uint16_t* dst = p + i * J;
    for (int h = 0; h < heads; ++h) {
      float v[NT];
      float mx = -INFINITY;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        v[t] = src[(long long)(h * NT + t) * n];
        mx = fmaxf(mx, v[t]);
      }
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        v[t] = ex2((v[t] - mx) * L2E);
        sum += v[t];
      }
      const float inv = 1.0f / sum;
      if constexpr (NT % 2 == 0) {   // heads * NT * 2 bytes per output row: 4-byte aligned pairs
#pragma unroll
        for (int t = 0; t < NT; t += 2)
          *reinterpret_cast<uint32_t*>(dst + h * NT + t) = pack2<BF16>(v[t] * inv, v[t + 1] * inv);
      } else {
#pragma unroll
        for (int t = 0; t < NT; ++t) dst[h * NT + t] = pack1<BF16>(v[t] * inv);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Y[g] (J x C) = P[g] (J x n) X[g] (n x C): one CTA per (query g, 128-channel slice). X tiles of 64 tokens x 128 channels
// (two 128-byte-swizzled TMA boxes) and the matching 48 x 64 tile of P stream through a 4-deep mbarrier ring; 8 consumer
// warps own 16 channels each: A fragments (P, row-major) by ldmatrix, B fragments (X, token-major = "V" of a flash
// kernel) by ldmatrix.trans, mma.sync m16n8k16 with fp32 accumulators (3 row tiles x 2 column tiles per warp).
// HBM-bound: X is read exactly once (738 MB at G = 128), P is re-read from L2 by the 11 slice CTAs of a query.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kWsStages = 4, kWsTok = 64, kWsJ = 48, kWsSlice = 128;
constexpr int kWsXBytes = kWsTok * 128;                    // one [64 tokens x 64 channels] box
constexpr int kWsPBytes = kWsJ * 128;                      // [48 rows x 64 tokens]
constexpr int kWsStageBytes = 2 * kWsXBytes + kWsPBytes;   // 22528
constexpr int kWsThreads = 288;                            // 8 consumer warps + the TMA producer warp
constexpr int kWsSmem = kWsStages * kWsStageBytes + 1024;

L4P_DEVICE void ws_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
L4P_DEVICE void ws_ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool BF16>
L4P_DEVICE void ws_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <bool BF16>
__global__ void __launch_bounds__(kWsThreads, 2)
token_weighted_sum_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmX,
                          uint16_t* __restrict__ y, int J, int n, int C, int nslices) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kWsStages], bar_empty[kWsStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int g = blockIdx.x / nslices, slice = blockIdx.x - g * nslices;
  const int nsteps = n / kWsTok;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kWsStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 8);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();  // P and X are outputs of the preceding kernels
  if (warp == 8) {
    if (lane == 0) {
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % kWsStages;
        if (it >= kWsStages) mbar_wait(smem_u32(&bar_empty[s]), ((uint32_t)(it / kWsStages) - 1u) & 1u);
        const uint32_t fb = smem_u32(&bar_full[s]);
        const uint32_t st = smem_base + (uint32_t)s * kWsStageBytes;
        mbar_expect_tx(fb, kWsStageBytes);
        tma_load_2d(st, &tmX, fb, slice * kWsSlice, g * n + it * kWsTok);
        tma_load_2d(st + kWsXBytes, &tmX, fb, slice * kWsSlice + 64, g * n + it * kWsTok);
        tma_load_2d(st + 2 * kWsXBytes, &tmP, fb, it * kWsTok, g * J);
      }
    }
    return;
  }
  float acc[3][2][4];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  // ldmatrix lane roles (see the header comment): B: matrix id = lane / 8 -> (token half, column tile); A: row = lane % 16,
  // token half = lane / 16
  const int bid = lane >> 3;
  const uint32_t b_row = (uint32_t)((bid & 1) * 8 + (lane & 7));
  const uint32_t b_chunk = (uint32_t)((warp & 3) * 2 + (bid >> 1));
  const uint32_t b_box = (uint32_t)(warp >> 2) * kWsXBytes;
  const uint32_t a_row = (uint32_t)(lane & 15), a_half = (uint32_t)(lane >> 4);
  for (int it = 0; it < nsteps; ++it) {
    const int s = it % kWsStages;
    mbar_wait(smem_u32(&bar_full[s]), (uint32_t)(it / kWsStages) & 1u);
    const uint32_t st = smem_base + (uint32_t)s * kWsStageBytes;
    const uint32_t sx = st + b_box, sp = st + 2 * kWsXBytes;
#pragma unroll
    for (int k16 = 0; k16 < kWsTok / 16; ++k16) {
      uint32_t b00, b01, b10, b11;
      {
        const uint32_t row = (uint32_t)k16 * 16u + b_row;
        ws_ldsm_x4_trans(sx + row * 128u + ((b_chunk ^ (row & 7u)) << 4), b00, b01, b10, b11);
      }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        uint32_t a0, a1, a2, a3;
        const uint32_t row = (uint32_t)mt * 16u + a_row;
        const uint32_t chunk = (uint32_t)k16 * 2u + a_half;
        ws_ldsm_x4(sp + row * 128u + ((chunk ^ (row & 7u)) << 4), a0, a1, a2, a3);
        ws_mma<BF16>(acc[mt][0], a0, a1, a2, a3, b00, b01);
        ws_mma<BF16>(acc[mt][1], a0, a1, a2, a3, b10, b11);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_empty[s]));
  }
  // accumulator layout: c0,c1 -> row lane/4, columns 2*(lane%4)+{0,1}; c2,c3 -> row + 8
  const int col0 = slice * kWsSlice + warp * 16 + 2 * (lane & 3);
#pragma unroll
  for (int mt = 0; mt < 3; ++mt) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int j = mt * 16 + (lane >> 2) + hh * 8;
      if (j < J) {
        uint16_t* dst = y + ((long long)g * J + j) * C;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int col = col0 + nt * 8;
          if (col < C) *reinterpret_cast<uint32_t*>(dst + col) = pack2<BF16>(acc[mt][nt][2 * hh], acc[mt][nt][2 * hh + 1]);
        }
      }
    }
  }
}

}  // namespace l4p

using namespace l4p;

extern "C" int l4p_head_expand(const float* q, void* out16, int64_t G, int nt, int heads, int hd, float scale, int bf16,
                               void* stream) {
  L4P_REQUIRE(q && out16, L4P_ERR_ARG, "l4p_head_expand: null pointer");
  L4P_REQUIRE(G >= 0 && nt > 0 && heads > 0 && hd > 0 && hd % 8 == 0, L4P_ERR_SHAPE, "l4p_head_expand: G=%lld nt=%d heads=%d hd=%d",
              (long long)G, nt, heads, hd);
  if (G == 0) return L4P_OK;
  const long long total8 = (long long)G * heads * nt * (heads * hd / 8);
  long long grid = (total8 + 255) / 256;
  if (grid > 16ll * host_num_sms()) grid = 16ll * host_num_sms();
  if (bf16)
    L4P_CHECK_CUDA(launch_pdl(head_expand_kernel<true>, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, q, (uint16_t*)out16, total8, nt, heads, hd, scale));
  else
    L4P_CHECK_CUDA(launch_pdl(head_expand_kernel<false>, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, q, (uint16_t*)out16, total8, nt, heads, hd, scale));
  return L4P_OK;
}

extern "C" int l4p_head_diag_gather(const float* z, float* out, int64_t G, int nt, int heads, int hd, void* stream) {
  L4P_REQUIRE(z && out, L4P_ERR_ARG, "l4p_head_diag_gather: null pointer");
  L4P_REQUIRE(G >= 0 && nt > 0 && heads > 0 && hd > 0 && hd % 4 == 0, L4P_ERR_SHAPE, "l4p_head_diag_gather: G=%lld nt=%d heads=%d hd=%d",
              (long long)G, nt, heads, hd);
  if (G == 0) return L4P_OK;
  const long long total4 = (long long)G * nt * (heads * hd / 4);
  long long grid = (total4 + 255) / 256;
  if (grid > 16ll * host_num_sms()) grid = 16ll * host_num_sms();
  L4P_CHECK_CUDA(launch_pdl(head_diag_gather_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, z, out, total4, nt, heads, hd));
  return L4P_OK;
}

extern "C" int l4p_row_softmax16(const float* s, void* p16, int64_t rows, int n, int bf16, void* stream) {
  L4P_REQUIRE(s && p16, L4P_ERR_ARG, "l4p_row_softmax16: null pointer");
  L4P_REQUIRE(rows >= 0 && n > 0 && n % 4 == 0 && n <= kRsVec * 128, L4P_ERR_SHAPE, "l4p_row_softmax16: n=%d (multiple of 4, <= %d)", n,
              kRsVec * 128);
  if (rows == 0) return L4P_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (bf16)
    L4P_CHECK_CUDA(launch_pdl(row_softmax16_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, s, (uint16_t*)p16, (long long)rows, n));
  else
    L4P_CHECK_CUDA(launch_pdl(row_softmax16_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, s, (uint16_t*)p16, (long long)rows, n));
  return L4P_OK;
}

extern "C" int l4p_token_weighted_sum(const void* p16, const void* x16, void* y16, int64_t G, int J, int n, int C, int bf16,
                                      void* stream) {
  L4P_REQUIRE(p16 && x16 && y16, L4P_ERR_ARG, "l4p_token_weighted_sum: null pointer");
  L4P_REQUIRE(G >= 0 && J > 0 && J <= kWsJ && n > 0 && n % kWsTok == 0 && C > 0 && C % 16 == 0, L4P_ERR_SHAPE,
              "l4p_token_weighted_sum: J=%d (<= %d) n=%d (multiple of %d) C=%d (multiple of 16)", J, kWsJ, n, kWsTok, C);
  L4P_REQUIRE(G * (int64_t)n < (1ll << 31) && G * (int64_t)J < (1ll << 31), L4P_ERR_SHAPE, "l4p_token_weighted_sum: G=%lld too large", (long long)G);
  if (G == 0) return L4P_OK;
  CUtensorMap tmP, tmX;
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)n, (uint64_t)G * J};
    const uint64_t strides[1] = {(uint64_t)n * 2};
    const uint32_t box[2] = {(uint32_t)kWsTok, (uint32_t)kWsJ};
    rc = host_make_tmap_16b(&tmP, p16, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)G * n};
    const uint64_t strides[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {64, (uint32_t)kWsTok};
    rc = host_make_tmap_16b(&tmX, x16, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  }
  const int nslices = (C + kWsSlice - 1) / kWsSlice;
  typedef void (*KFn)(const CUtensorMap, const CUtensorMap, uint16_t*, int, int, int, int);
  KFn kfn = bf16 ? token_weighted_sum_kernel<true> : token_weighted_sum_kernel<false>;
  L4P_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmem));
  L4P_CHECK_CUDA(launch_pdl(kfn, dim3((unsigned)(G * nslices)), dim3(kWsThreads), (size_t)kWsSmem, (cudaStream_t)stream, tmP, tmX,
                            (uint16_t*)y16, J, n, C, nslices));
  return L4P_OK;
}

extern "C" int l4p_group_softmax_t16(const float* s, void* p16, int64_t G, int heads, int nt, int n, int bf16, void* stream) {
  L4P_REQUIRE(s && p16, L4P_ERR_ARG, "l4p_group_softmax_t16: null pointer");
  L4P_REQUIRE(G >= 0 && heads > 0 && n > 0 && nt >= 1 && nt <= 8, L4P_ERR_SHAPE, "l4p_group_softmax_t16: G=%lld heads=%d nt=%d n=%d (nt <= 8)",
              (long long)G, heads, nt, n);
  if (G == 0) return L4P_OK;
  const long long total = (long long)G * n;
  long long grid = (total + 255) / 256;
  if (grid > 32ll * host_num_sms()) grid = 32ll * host_num_sms();
  typedef void (*KFn)(const float*, uint16_t*, long long, int, int);
  static const KFn table[2][8] = {
      {group_softmax_t16_kernel<false, 1>, group_softmax_t16_kernel<false, 2>, group_softmax_t16_kernel<false, 3>, group_softmax_t16_kernel<false, 4>,
       group_softmax_t16_kernel<false, 5>, group_softmax_t16_kernel<false, 6>, group_softmax_t16_kernel<false, 7>, group_softmax_t16_kernel<false, 8>},
      {group_softmax_t16_kernel<true, 1>, group_softmax_t16_kernel<true, 2>, group_softmax_t16_kernel<true, 3>, group_softmax_t16_kernel<true, 4>,
       group_softmax_t16_kernel<true, 5>, group_softmax_t16_kernel<true, 6>, group_softmax_t16_kernel<true, 7>, group_softmax_t16_kernel<true, 8>}};
  L4P_CHECK_CUDA(launch_pdl(table[bf16 ? 1 : 0][nt - 1], dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, s, (uint16_t*)p16, total, n, heads));
  return L4P_OK;
}
