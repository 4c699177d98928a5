// EXPERIMENTAL (opt-in, L4P_IMGATT_TC=1; not yet validated on hardware - see DESIGN.md section 7): tensor-core
// formulation of the image -> token cross attention of SAM's two-way transformer (sam/transformer.py:179-184).
//
// The CUDA-core kernel (track.cu: image_attention_kernel) is FMA-bound: 2 x nk x 88 FMAs per (row, head), 123 us of
// FMA-pipe time per launch at 262 144 rows. Here a unit = (128 rows, one head):
//     S[128 x 16]  = Q_h[128 x 96] K_h^T          6 UMMAs  (M128 N16 K16; keys padded 6 -> 16, channels 88 -> 96)
//     P            = softmax over the nk valid columns (one row per thread), padded to 64 keys, 16-bit, smem
//     O[128 x 96]  = P[128 x 64] V_h[64 x 96]     4 UMMAs  (M128 N96 K16)
// ~10 UMMAs x ~95 cycles per unit instead of 1056 FMAs per thread. K/V are tiny (G x nk x 704 fp32): a prep kernel
// converts them once per launch into the padded 16-bit operand layouts (Kp [G*H*16, 96] pre-scaled by scale*log2e,
// Vt [G*H*96, 64]) inside a caller-provided workspace.
//
//   warps 0..3  softmax + epilogue (row per thread)      warp 4  TMA producer      warp 5  UMMA issuer      warp 6  TMEM
#include "common.cuh"

namespace l4p {

constexpr int kItD = 88, kItDPad = 96, kItKeys = 16, kItKeysPad = 64, kItRows = 128;
constexpr int kItQBytes = kItRows * kItDPad * 2;        // 24576: 3 chunks x (128 rows x 64 B), SWIZZLE_64B
constexpr int kItVBytes = kItDPad * kItKeysPad * 2;     // 12288: 96 rows x 128 B, SWIZZLE_128B
constexpr int kItKBytes = kItKeys * kItDPad * 2;        //  3072: 3 chunks x (16 rows x 64 B), SWIZZLE_64B
constexpr int kItStageBytes = kItQBytes + kItVBytes + kItKBytes;  // 39936 = 39 KiB (every part 1 KiB aligned)
constexpr int kItPBytes = kItRows * 128;                // 16384: P tile, 128 rows x (64 keys x 2 B), SWIZZLE_128B
constexpr int kItStages = 2;
constexpr int kItSmem = kItStages * (kItStageBytes + kItPBytes) + 1024;
constexpr int kItThreads = 256;
constexpr uint32_t kItColS = 0, kItColO = 32, kItColStage = 128;  // TMEM columns per stage: S [0,16), O [32,128)

// k, v fp32 [G, nk, H*D] -> Kp 16-bit [G*H*16, 96] (scaled, zero padded), Vt 16-bit [G*H*96, 64] (transposed, zero padded)
template <bool BF16>
__global__ void imgatt_prep_kernel(const float* __restrict__ kf, const float* __restrict__ vf, uint16_t* __restrict__ kp,
                                   uint16_t* __restrict__ vt, int G, int nk, int H, float kscale) {
  const int ld = H * kItD;
  const long long nK = (long long)G * H * kItKeys * kItDPad, nV = (long long)G * H * kItDPad * kItKeysPad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nK + nV; i += (long long)gridDim.x * blockDim.x) {
    if (i < nK) {
      const int c = (int)(i % kItDPad);
      const int key = (int)((i / kItDPad) % kItKeys);
      const long long gh = i / (kItDPad * kItKeys);
      const int h = (int)(gh % H);
      const long long g = gh / H;
      const float x = (key < nk && c < kItD) ? kf[(g * nk + key) * ld + h * kItD + c] * kscale : 0.f;
      kp[i] = pack1<BF16>(x);
    } else {
      const long long j = i - nK;
      const int key = (int)(j % kItKeysPad);
      const int c = (int)((j / kItKeysPad) % kItDPad);
      const long long gh = j / (kItKeysPad * kItDPad);
      const int h = (int)(gh % H);
      const long long g = gh / H;
      const float x = (key < nk && c < kItD) ? vf[(g * nk + key) * ld + h * kItD + c] : 0.f;
      vt[j] = pack1<BF16>(x);
    }
  }
}

struct ImgAttParams {
  uint16_t* out;
  int Np, nk, H, tiles;  // tiles = G * Np / 128
};

template <bool BF16>
__global__ void __launch_bounds__(kItThreads, 1)
image_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const ImgAttParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kItStages], bar_empty[kItStages], bar_sfull[kItStages], bar_pfull[kItStages],
      bar_ofull[kItStages], bar_ofree[kItStages];
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sStage = smem_base;                                 // [stage]: Q | V | K
  const uint32_t sP = smem_base + kItStages * kItStageBytes;         // [stage]: P
  const int units = p.tiles * p.H;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int s = 0; s < kItStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
      mbar_init(smem_u32(&bar_sfull[s]), 1);
      mbar_init(smem_u32(&bar_pfull[s]), 128);
      mbar_init(smem_u32(&bar_ofull[s]), 1);
      mbar_init(smem_u32(&bar_ofree[s]), 128);
    }
    fence_mbar_init();
  }
  // the P tiles are zero outside their first 16-byte chunk (keys 8..63 do not exist): clear them once
  for (int i = threadIdx.x; i < kItStages * kItPBytes / 16; i += kItThreads)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sP + (uint32_t)i * 16u), "r"(0u) : "memory");
  fence_proxy_async();
  if (warp == 6) tmem_alloc(smem_u32(&tmem_base_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int u = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++u) {
        const int s = u & 1;
        const int tile = unit / p.H, h = unit - tile * p.H;
        const int g = (int)(((long long)tile * kItRows) / p.Np);
        mbar_wait(smem_u32(&bar_empty[s]), (((uint32_t)u >> 1) & 1u) ^ 1u);
        const uint32_t fb = smem_u32(&bar_full[s]);
        const uint32_t sq = sStage + s * kItStageBytes, sv = sq + kItQBytes, sk = sv + kItVBytes;
        mbar_expect_tx(fb, kItStageBytes);
        for (int c = 0; c < 3; ++c) tma_load_2d(sq + c * (kItRows * 64), &tmQ, fb, h * kItD + c * 32, tile * kItRows);
        tma_load_2d(sv, &tmV, fb, 0, (g * p.H + h) * kItDPad);
        for (int c = 0; c < 3; ++c) tma_load_2d(sk + c * (kItKeys * 64), &tmK, fb, c * 32, (g * p.H + h) * kItKeys);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ UMMA issuer
    const bool leader = elect_one();
    const uint32_t idesc_s = umma_idesc_f16(BF16, kItRows, kItKeys);
    const uint32_t idesc_o = umma_idesc_f16(BF16, kItRows, kItDPad);
    constexpr uint32_t hi64 = umma_desc_hi(64, 4), hi128 = umma_desc_hi(128, 2);
    int u = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++u) {
      const int s = u & 1;
      const uint32_t ph = ((uint32_t)u >> 1) & 1u;
      const uint32_t sq = sStage + s * kItStageBytes, sv = sq + kItQBytes, sk = sv + kItVBytes;
      const uint32_t tS = tmem_base + (uint32_t)s * kItColStage + kItColS, tO = tmem_base + (uint32_t)s * kItColStage + kItColO;
      mbar_wait(smem_u32(&bar_full[s]), ph);
      mbar_wait(smem_u32(&bar_ofree[s]), ph ^ 1u);  // the epilogue of unit u-2 has drained this TMEM stage
      tc_fence_after();
      if (leader) {
        const uint32_t q_lo = umma_desc_lo(sq), k_lo = umma_desc_lo(sk);
#pragma unroll
        for (int kk = 0; kk < kItDPad / 16; ++kk) {
          const uint32_t qoff = ((uint32_t)(kk >> 1) * (kItRows * 64) + (uint32_t)(kk & 1) * 32) >> 4;
          const uint32_t koff = ((uint32_t)(kk >> 1) * (kItKeys * 64) + (uint32_t)(kk & 1) * 32) >> 4;
          umma_ss(tS, umma_desc_make(q_lo + qoff, hi64), umma_desc_make(k_lo + koff, hi64), idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar_sfull[s]));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar_pfull[s]), ph);
      tc_fence_after();
      if (leader) {
        const uint32_t p_lo = umma_desc_lo(sP + s * kItPBytes), v_lo = umma_desc_lo(sv);
#pragma unroll
        for (int kk = 0; kk < kItKeysPad / 16; ++kk)
          umma_ss(tO, umma_desc_make(p_lo + 2 * kk, hi128), umma_desc_make(v_lo + 2 * kk, hi128), idesc_o, kk != 0 ? 1u : 0u);
        umma_commit(smem_u32(&bar_ofull[s]));
        umma_commit(smem_u32(&bar_empty[s]));  // Q / K / V / P of this stage are consumed
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ softmax + epilogue (row per thread)
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    int u = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++u) {
      const int s = u & 1;
      const uint32_t ph = ((uint32_t)u >> 1) & 1u;
      const int tile = unit / p.H, h = unit - tile * p.H;
      const uint32_t tS = tmem_base + lane_addr + (uint32_t)s * kItColStage + kItColS;
      const uint32_t tO = tmem_base + lane_addr + (uint32_t)s * kItColStage + kItColO;
      mbar_wait(smem_u32(&bar_sfull[s]), ph);
      tc_fence_after();
      uint32_t sraw[16];
      tmem_ld16(tS, sraw);
      tmem_ld_wait();
      float pj[8];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < p.nk) mx = fmaxf(mx, __uint_as_float(sraw[j]));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        pj[j] = j < p.nk ? ex2(__uint_as_float(sraw[j]) - mx) : 0.f;   // K was pre-scaled by scale * log2(e)
        sum += pj[j];
      }
      const float inv = 1.f / sum;
      // P row: keys 0..7 in the first 16-byte chunk of the 128-byte swizzled row, everything else stays zero.
      // The stage's P buffer is free: PV of unit u-2 completed before ofull(u-2), which this thread waited for.
      const uint32_t paddr = sP + s * kItPBytes + (uint32_t)r * 128u + ((0u ^ ((uint32_t)r & 7u)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(paddr), "r"(pack2<BF16>(pj[0] * inv, pj[1] * inv)),
                   "r"(pack2<BF16>(pj[2] * inv, pj[3] * inv)), "r"(pack2<BF16>(pj[4] * inv, pj[5] * inv)),
                   "r"(pack2<BF16>(pj[6] * inv, pj[7] * inv))
                   : "memory");
      fence_proxy_async();
      mbar_arrive(smem_u32(&bar_pfull[s]));

      mbar_wait(smem_u32(&bar_ofull[s]), ph);
      tc_fence_after();
      const long long row = (long long)tile * kItRows + r;
      uint16_t* dst = p.out + row * ((long long)p.H * kItD) + h * kItD;
#pragma unroll
      for (int cc = 0; cc < kItDPad; cc += 32) {
        uint32_t o[32];
        tmem_ld32(tO + cc, o);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int col = cc + g8 * 8;
          if (col < kItD)
            *reinterpret_cast<uint4*>(dst + col) =
                make_uint4(pack2<BF16>(__uint_as_float(o[g8 * 8 + 0]), __uint_as_float(o[g8 * 8 + 1])),
                           pack2<BF16>(__uint_as_float(o[g8 * 8 + 2]), __uint_as_float(o[g8 * 8 + 3])),
                           pack2<BF16>(__uint_as_float(o[g8 * 8 + 4]), __uint_as_float(o[g8 * 8 + 5])),
                           pack2<BF16>(__uint_as_float(o[g8 * 8 + 6]), __uint_as_float(o[g8 * 8 + 7])));
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_ofree[s]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace l4p

using namespace l4p;

extern "C" int64_t l4p_image_attention_tc_workspace_bytes(int G, int H) {
  return (int64_t)G * H * (kItKeys * kItDPad + kItDPad * kItKeysPad) * 2;
}

extern "C" int l4p_image_attention_tc(const void* q16, const float* k, const float* v, void* out16, void* ws, int64_t ws_bytes,
                                      int G, int Np, int nk, int H, int d, float scale, int bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  L4P_REQUIRE(q16 && k && v && out16 && ws, L4P_ERR_ARG, "l4p_image_attention_tc: null pointer");
  L4P_REQUIRE(d == kItD && nk >= 1 && nk <= 8 && G > 0 && H > 0 && Np % kItRows == 0, L4P_ERR_SHAPE,
              "l4p_image_attention_tc: d=%d (88) nk=%d (<=8) Np=%d (multiple of 128)", d, nk, Np);
  L4P_REQUIRE(ws_bytes >= l4p_image_attention_tc_workspace_bytes(G, H), L4P_ERR_ARG, "l4p_image_attention_tc: workspace too small");
  uint16_t* kp = (uint16_t*)ws;
  uint16_t* vt = kp + (size_t)G * H * kItKeys * kItDPad;
  const float kscale = scale * 1.4426950408889634f;
  if (bf16) imgatt_prep_kernel<true><<<64, 256, 0, stream>>>(k, v, kp, vt, G, nk, H, kscale);
  else imgatt_prep_kernel<false><<<64, 256, 0, stream>>>(k, v, kp, vt, G, nk, H, kscale);
  L4P_CHECK_CUDA(cudaGetLastError());

  CUtensorMap tmQ, tmK, tmV;
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)H * kItD, (uint64_t)G * Np};
    const uint64_t strides[1] = {(uint64_t)H * kItD * 2};
    const uint32_t box[2] = {32, (uint32_t)kItRows};
    rc = host_make_tmap_16b(&tmQ, q16, 2, dims, strides, box, 64);
    if (rc != L4P_OK) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)kItDPad, (uint64_t)G * H * kItKeys};
    const uint64_t strides[1] = {(uint64_t)kItDPad * 2};
    const uint32_t box[2] = {32, (uint32_t)kItKeys};
    rc = host_make_tmap_16b(&tmK, kp, 2, dims, strides, box, 64);
    if (rc != L4P_OK) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)kItKeysPad, (uint64_t)G * H * kItDPad};
    const uint64_t strides[1] = {(uint64_t)kItKeysPad * 2};
    const uint32_t box[2] = {(uint32_t)kItKeysPad, (uint32_t)kItDPad};
    rc = host_make_tmap_16b(&tmV, vt, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  }
  ImgAttParams p;
  p.out = (uint16_t*)out16;
  p.Np = Np; p.nk = nk; p.H = H;
  p.tiles = (int)(((long long)G * Np) / kItRows);
  const long long units = (long long)p.tiles * H;
  int grid = host_num_sms();
  if (grid > units) grid = (int)units;
  auto kfn = bf16 ? image_attention_tc_kernel<true> : image_attention_tc_kernel<false>;
  L4P_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kItSmem));
  L4P_CHECK_CUDA(launch_pdl(kfn, dim3(grid), dim3(kItThreads), (size_t)kItSmem, stream, tmQ, tmK, tmV, p));
  return L4P_OK;
}
