// K4 (experimental CTA-pair variant, opt-in with L4P_ATT_PAIR=1; NOT yet run on hardware - written at the end of round 1
// when the GPU budget was spent): the fused attention kernel of attention.cu with
// tcgen05.mma.cta_group::2. A cluster of two CTAs owns 512 query rows of one (batch, head): each CTA keeps two 128-row
// query tiles (as attention.cu), but every key / value block is staged HALF per CTA - 64 of the 128 keys of K_j, 48 of the
// 96 rows of V^T_j - and the leader CTA issues M=256 UMMAs that read both CTAs' shared memory and write each CTA's own TMEM.
// Per SM and key block that is 24 KiB instead of 48 KiB of TMA writes and 128 KiB instead of 176 KiB of UMMA operand reads
// through the 128 B/clk shared-memory port (the measured limiter of attention.cu), and half the single-thread UMMA issue
// work per SM. The softmax warpgroups are attention.cu's, unchanged except that their "S consumed" / "P ready" arrivals go
// to the leader CTA's barriers (2 x 128 arrivals each).
//
//   barrier            lives in     arrivals
//   bar_q              leader       1 (leader's expect_tx) + bytes of both CTAs' Q tiles
//   bar_kfull/vfull[s] leader       1 (leader's expect_tx) + bytes of both halves
//   bar_kempty/vempty  each CTA     multicast tcgen05.commit
//   bar_sfull[t]       each CTA     multicast tcgen05.commit (S_t complete in both TMEMs)
//   bar_sfree[t]       leader       256 (every softmax thread of tile t in both CTAs holds S_t in registers)
//   bar_pfull[t]       leader       256 (P_t of both CTAs written, O_t rescaled)
//   bar_pvdone[t]      each CTA     multicast tcgen05.commit
//
// Every wait is the bounded mbar_wait of common.cuh: a protocol bug traps instead of hanging the GPU.
#include <cstdlib>

#include "common.cuh"

namespace l4p {
namespace attpair {

constexpr int kAttThreads = 384;
constexpr int kDPad = 96;
constexpr int kTileM = 128;   // query rows per tile and CTA
constexpr int kTileN = 128;   // keys per block
constexpr int kQTileBytes = kTileM * kDPad * 2;        // 24576: 3 chunks x (128 rows x 64 B)
constexpr int kKHalfBytes = (kTileN / 2) * kDPad * 2;  // 12288: 3 chunks x (64 keys x 64 B)
constexpr int kVHalfBytes = (kDPad / 2) * kTileN * 2;  // 12288: 2 chunks x (48 rows x 128 B)
constexpr int kKChunk = (kTileN / 2) * 64;             // 4096 B between the 32-element chunks of a K half
constexpr int kVChunk = (kDPad / 2) * 128;             // 6144 B between the 64-key chunks of a V^T half
constexpr int kPBytes = kTileM * kTileN * 2;           // 32768: tile 1's P, 2 chunks x (128 rows x 128 B)
constexpr int kKS = 4, kVS = 4;
constexpr int kAttSmem = 2 * kQTileBytes + kKS * kKHalfBytes + kVS * kVHalfBytes + kPBytes + 1024;
constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 352, kColP0 = 448;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
#ifndef L4P_ATT_P_ALIAS
#define L4P_ATT_P_ALIAS 0
#endif
constexpr bool kPAlias = L4P_ATT_P_ALIAS != 0;  // as in attention.cu: both tiles' P overwrite S_t[0,64) in TMEM, PV always TS-mode

struct AttParams {
  uint16_t* out;
  int B, H, N, head_dim;
  float scale_log2;  // scale * log2(e)
};

// D[tmem, both CTAs] (+)= A[tmem, each CTA's own 128 lanes] * B[smem, N/2 rows per CTA]
L4P_DEVICE void umma2_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 2^x for a pair on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], cubic minimax for
// 2^f (max rel err 7.6e-5, well below the 16-bit rounding of P), exponent patched in with an integer add.
L4P_DEVICE uint64_t exp2_poly2(uint64_t t2) {
  float a, b;
  upk2(t2, a, b);
  a = fmaxf(a, -125.0f);
  b = fmaxf(b, -125.0f);
  t2 = pk2(a, b);
  const uint64_t magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f);
  const uint64_t mone = pk2(-1.0f, -1.0f);
  const uint64_t xf = add2(t2, magic);   // low mantissa bits = round(x)
  const uint64_t n2 = add2(xf, nmagic);
  const uint64_t f2 = fma2(n2, mone, t2);
  uint64_t p2 = fma2(pk2(0.05520550534129143f, 0.05520550534129143f), f2, pk2(0.24261397123336792f, 0.24261397123336792f));
  p2 = fma2(p2, f2, pk2(0.6932547688484192f, 0.6932547688484192f));
  p2 = fma2(p2, f2, pk2(0.9999276995658875f, 0.9999276995658875f));
  float pa, pb, xa, xb;
  upk2(p2, pa, pb);
  upk2(xf, xa, xb);
  pa = __int_as_float(__float_as_int(pa) + (__float_as_int(xa) << 23));
  pb = __int_as_float(__float_as_int(pb) + (__float_as_int(xb) << 23));
  return pk2(pa, pb);
}

template <bool BF16, int POLY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAttThreads, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q;
  __shared__ __align__(8) uint64_t bar_kfull[kKS], bar_kempty[kKS];
  __shared__ __align__(8) uint64_t bar_vfull[kVS], bar_vempty[kVS];
  __shared__ __align__(8) uint64_t bar_sfull[2], bar_sfree[2], bar_pfull[2], bar_pvdone[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + 2 * kQTileBytes;
  const uint32_t sV = sK + kKS * kKHalfBytes;
  const uint32_t sP = sV + kVS * kVHalfBytes;

  const int clusters_per_head = p.N / (4 * kTileM);
  const int cluster = blockIdx.x >> 1;
  const int cl = cluster % clusters_per_head;
  const int bh = cluster / clusters_per_head;  // b * H + h
  const int nblk = p.N / kTileN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&bar_q), 1);
    for (int s = 0; s < kKS; ++s) { mbar_init(smem_u32(&bar_kfull[s]), 1); mbar_init(smem_u32(&bar_kempty[s]), 1); }
    for (int s = 0; s < kVS; ++s) { mbar_init(smem_u32(&bar_vfull[s]), 1); mbar_init(smem_u32(&bar_vempty[s]), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bar_sfull[t]), 1);
      mbar_init(smem_u32(&bar_sfree[t]), 256);
      mbar_init(smem_u32(&bar_pfull[t]), 256);
      mbar_init(smem_u32(&bar_pvdone[t]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc2(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit / peer TMA signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();  // Q/K/V are the previous kernel's outputs

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer (both CTAs: own Q tiles, own halves)
      const int q_row0 = bh * p.N + cl * 4 * kTileM;  // row in the [B*H*N, 96] view
      const uint32_t qb = mapa_shared(smem_u32(&bar_q), 0);
      if (is_leader) mbar_expect_tx(smem_u32(&bar_q), 4 * kQTileBytes);
      for (int t = 0; t < 2; ++t)
        for (int c = 0; c < 3; ++c)
          tma2_load_2d(sQ + t * kQTileBytes + c * (kTileM * 64), &tmQ, qb, c * 32, q_row0 + (t * 2 + (int)rank) * kTileM);
      for (int j = 0; j < nblk; ++j) {
        {
          const int s = j % kKS;
          mbar_wait(smem_u32(&bar_kempty[s]), (((uint32_t)(j / kKS)) & 1u) ^ 1u);
          const uint32_t fb = mapa_shared(smem_u32(&bar_kfull[s]), 0);
          if (is_leader) mbar_expect_tx(smem_u32(&bar_kfull[s]), 2 * kKHalfBytes);
          for (int c = 0; c < 3; ++c)
            tma2_load_2d(sK + s * kKHalfBytes + c * kKChunk, &tmK, fb, c * 32, bh * p.N + j * kTileN + (int)rank * (kTileN / 2));
        }
        {
          const int s = j % kVS;
          mbar_wait(smem_u32(&bar_vempty[s]), (((uint32_t)(j / kVS)) & 1u) ^ 1u);
          const uint32_t fb = mapa_shared(smem_u32(&bar_vfull[s]), 0);
          if (is_leader) mbar_expect_tx(smem_u32(&bar_vfull[s]), 2 * kVHalfBytes);
          for (int c = 0; c < 2; ++c)
            tma2_load_2d(sV + s * kVHalfBytes + c * kVChunk, &tmV, fb, j * kTileN + c * 64, bh * kDPad + (int)rank * (kDPad / 2));
        }
      }
    } else if (warp == 1 && is_leader) {
      // ---------------------------------------------------------------- UMMA issuer (leader CTA only, M = 256)
      const bool leader = elect_one();
      const uint32_t idesc_s = umma_idesc_f16(BF16, 2 * kTileM, kTileN);
      const uint32_t idesc_o = umma_idesc_f16(BF16, 2 * kTileM, kDPad);
      constexpr uint32_t hi64 = umma_desc_hi(64, 4), hi128 = umma_desc_hi(128, 2);
      const uint32_t q_lo = umma_desc_lo(sQ), k_lo = umma_desc_lo(sK), p_lo = umma_desc_lo(sP), v_lo = umma_desc_lo(sV);
      auto issue_s = [&](const int t, const int s) {
        const uint32_t d = tmem_base + (t == 0 ? kColS0 : kColS1);
        const uint32_t qa = q_lo + (uint32_t)t * (kQTileBytes >> 4), ka = k_lo + (uint32_t)s * (kKHalfBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < kDPad / 16; ++kk) {
          const uint32_t qoff = ((uint32_t)(kk >> 1) * (kTileM * 64) + (uint32_t)(kk & 1) * 32) >> 4;
          const uint32_t koff = ((uint32_t)(kk >> 1) * kKChunk + (uint32_t)(kk & 1) * 32) >> 4;
          umma2_ss(d, umma_desc_make(qa + qoff, hi64), umma_desc_make(ka + koff, hi64), idesc_s, kk != 0 ? 1u : 0u);
        }
        umma2_commit_mc(smem_u32(&bar_sfull[t]), 3);
      };
      auto issue_pv = [&](const int t, const int s, const uint32_t acc) {
        const uint32_t d = tmem_base + (t == 0 ? kColO0 : kColO1);
        const uint32_t va = v_lo + (uint32_t)s * (kVHalfBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < kTileN / 16; ++kk) {
          const uint32_t o = ((uint32_t)(kk & 3) * 32) >> 4;
          const uint64_t vdesc = umma_desc_make(va + (uint32_t)(kk >> 2) * (kVChunk >> 4) + o, hi128);
          if (kPAlias)
            umma2_ts(d, tmem_base + (t == 0 ? kColS0 : kColS1) + (uint32_t)kk * 8u, vdesc, idesc_o, kk != 0 ? 1u : acc);
          else if (t == 0)  // P_0 is each CTA's TMEM A operand: 8 columns per K = 16 step
            umma2_ts(d, tmem_base + kColP0 + (uint32_t)kk * 8u, vdesc, idesc_o, kk != 0 ? 1u : acc);
          else
            umma2_ss(d, umma_desc_make(p_lo + (uint32_t)(kk >> 2) * ((kTileM * 128) >> 4) + o, hi128), vdesc, idesc_o,
                     kk != 0 ? 1u : acc);
        }
        umma2_commit_mc(smem_u32(&bar_pvdone[t]), 3);
      };

      mbar_wait(smem_u32(&bar_q), 0);
      mbar_wait(smem_u32(&bar_kfull[0]), 0);
      tc_fence_after();
      if (leader) {
        issue_s(0, 0);
        issue_s(1, 0);
        umma2_commit_mc(smem_u32(&bar_kempty[0]), 3);
      }
      __syncwarp();
      for (int j = 0; j < nblk; ++j) {
        const int jn = j + 1;
        const int sk = jn % kKS, sv = j % kVS;
#if L4P_ATT_P_ALIAS
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(smem_u32(&bar_pfull[t]), (uint32_t)j & 1u);  // P_t(j) sits in S_t's columns in both CTAs, O_t rescaled
          if (t == 0) mbar_wait(smem_u32(&bar_vfull[sv]), ((uint32_t)(j / kVS)) & 1u);
          if (jn < nblk && t == 0) mbar_wait(smem_u32(&bar_kfull[sk]), ((uint32_t)(jn / kKS)) & 1u);
          tc_fence_after();
          if (leader) {
            issue_pv(t, sv, j != 0 ? 1u : 0u);
            if (t == 1) umma2_commit_mc(smem_u32(&bar_vempty[sv]), 3);
            if (jn < nblk) {
              issue_s(t, sk);  // overwrites S_t / P_t(j) after PV_t(j) in pipe order
              if (t == 1) umma2_commit_mc(smem_u32(&bar_kempty[sk]), 3);
            }
          }
          __syncwarp();
        }
        continue;
#endif
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (jn < nblk) {
            if (t == 0) mbar_wait(smem_u32(&bar_kfull[sk]), ((uint32_t)(jn / kKS)) & 1u);
            mbar_wait(smem_u32(&bar_sfree[t]), (uint32_t)j & 1u);  // both CTAs' softmax t hold S_t(j) in registers
            tc_fence_after();
            if (leader) {
              issue_s(t, sk);
              if (t == 1) umma2_commit_mc(smem_u32(&bar_kempty[sk]), 3);
            }
            __syncwarp();
          }
          mbar_wait(smem_u32(&bar_pfull[t]), (uint32_t)j & 1u);  // P_t(j) ready in both CTAs (and O_t rescaled)
          if (t == 0) mbar_wait(smem_u32(&bar_vfull[sv]), ((uint32_t)(j / kVS)) & 1u);
          tc_fence_after();
          if (leader) {
            issue_pv(t, sv, j != 0 ? 1u : 0u);
            if (t == 1) umma2_commit_mc(smem_u32(&bar_vempty[sv]), 3);
          }
          __syncwarp();
        }
      }
      // the last commit (release of the last V slot, multicast) has no other waiter: wait for its local arrival so that no
      // multicast arrive is in flight towards the peer's shared memory when the cluster exits (tools/att_protocol_sim.py)
      mbar_wait(smem_u32(&bar_vempty[(nblk - 1) % kVS]), ((uint32_t)((nblk - 1) / kVS)) & 1u);
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int t = (warp - 4) >> 2;            // tile
    const int q4 = warp & 3;                   // TMEM lane quarter
    const int r = q4 * 32 + lane;              // row in tile
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + (t == 0 ? kColS0 : kColS1);
    const uint32_t tO = tmem_base + lane_addr + (t == 0 ? kColO0 : kColO1);
    const uint32_t pRow = sP + (uint32_t)r * 128u;  // only tile 1 stages P in shared memory
    const uint32_t swz = (uint32_t)(r & 7);
    const float c = p.scale_log2;
    const uint32_t sfree_leader = mapa_shared(smem_u32(&bar_sfree[t]), 0);
    const uint32_t pfull_leader = mapa_shared(smem_u32(&bar_pfull[t]), 0);

    float m_used = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(smem_u32(&bar_sfull[t]), (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t s[128];
      tmem_ld32(tS + 0, s + 0);
      tmem_ld32(tS + 32, s + 32);
      tmem_ld32(tS + 64, s + 64);
      tmem_ld32(tS + 96, s + 96);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_cluster(sfree_leader);

      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]),
            mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i < 128; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;

      if (j == 0) {
        m_used = mx;
      } else {
        const bool need = mx > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // rare: rescale the running output in TMEM (whole warp, each row with its own factor)
          const float m_new = fmaxf(m_used, mx);
          const float alpha = ex2(m_used - m_new);
          mbar_wait(smem_u32(&bar_pvdone[t]), (uint32_t)(j - 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int cc = 0; cc < kDPad; cc += 32) {
            uint32_t o[32];
            tmem_ld32(tO + cc, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + cc, o);
            tmem_st16(tO + cc + 16, o + 16);
          }
          tmem_st_wait();
          tc_fence_before();
          l *= alpha;
          m_used = m_new;
        }
      }
      // p = exp2(s*c - m_used), row sum in fp32, pack pairs in place
      // scale/subtract and the row sum run as packed f32x2; of every 16 scores, 6 take the polynomial exp2 on the
      // FMA pipe and 10 the MUFU, which balances the two pipes (MUFU alone is the measured bound of this kernel)
      const uint64_t c2 = pk2(c, c), nm2 = pk2(-m_used, -m_used);
      uint64_t lsum0 = pk2(0.f, 0.f), lsum1 = pk2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 128; i += 2) {
        const uint64_t t2 = fma2(pk2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, nm2);
        uint64_t p2;
        const int slot = (i >> 1) & 7;  // pair index inside a group of 16 scores
        if ((POLY == 3 && (slot == 1 || slot == 4 || slot == 6)) || (POLY == 2 && (slot == 1 || slot == 5)) ||
            (POLY == 1 && slot == 3)) {
          p2 = exp2_poly2(t2);
        } else {
          float a, b;
          upk2(t2, a, b);
          p2 = pk2(ex2(a), ex2(b));
        }
        if (i & 2) lsum1 = add2(lsum1, p2); else lsum0 = add2(lsum0, p2);
        float p0, p1;
        upk2(p2, p0, p1);
        s[i >> 1] = pack2<BF16>(p0, p1);
      }
      {
        float a0, a1, b0, b1;
        upk2(lsum0, a0, a1);
        upk2(lsum1, b0, b1);
        l += (a0 + a1) + (b0 + b1);
      }

      if (j > 0) mbar_wait(smem_u32(&bar_pvdone[t]), (uint32_t)(j - 1) & 1u);  // P_t smem is free again
      if (kPAlias || t == 0) {
        // tile 0: P goes straight back to TMEM (row = lane, two probabilities per 32-bit column) as the A operand of PV
        const uint32_t tP = kPAlias ? tS : tmem_base + lane_addr + kColP0;
        tmem_st16(tP + 0, s + 0);
        tmem_st16(tP + 16, s + 16);
        tmem_st16(tP + 32, s + 32);
        tmem_st16(tP + 48, s + 48);
        tmem_st_wait();
        tc_fence_before();
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          // 16-byte unit u: keys [8u, 8u+8); chunk = u/8; swizzled unit inside the 128-byte row
          const uint32_t addr = pRow + (uint32_t)(u >> 3) * (kTileM * 128) + ((((uint32_t)u & 7u) ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(s[4 * u]), "r"(s[4 * u + 1]),
                       "r"(s[4 * u + 2]), "r"(s[4 * u + 3])
                       : "memory");
        }
        fence_proxy_async();
      }
      mbar_arrive_cluster(pfull_leader);
    }

    // ---- epilogue: O_t / l -> global
    mbar_wait(smem_u32(&bar_pvdone[t]), (uint32_t)(nblk - 1) & 1u);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int b = bh / p.H, h = bh - b * p.H;
    const long long row = (long long)b * p.N + (long long)cl * 4 * kTileM + (t * 2 + (int)rank) * kTileM + r;
    uint16_t* dst = p.out + row * ((long long)p.H * p.head_dim) + (long long)h * p.head_dim;
#pragma unroll
    for (int cc = 0; cc < kDPad; cc += 32) {
      uint32_t o[32];
      tmem_ld32(tO + cc, o);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = cc + g * 8;
        if (col < p.head_dim) {  // head_dim is a multiple of 8
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[g * 8 + i]) * inv_l;
          *reinterpret_cast<uint4*>(dst + col) = make_uint4(pack2<BF16>(f[0], f[1]), pack2<BF16>(f[2], f[3]),
                                                            pack2<BF16>(f[4], f[5]), pack2<BF16>(f[6], f[7]));
        }
      }
    }
    tc_fence_before();
  }


  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer's UMMAs may still read this CTA's smem / TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace attpair

using namespace attpair;

// Host launcher, called by l4p_attention() when L4P_ATT_PAIR=1 and N is a multiple of 512.
int attention_pair_launch(const void* q, const void* k, const void* vt, void* out, int B, int H, int N, int head_dim,
                          float scale, int bf16, int poly, void* stream) {
  L4P_REQUIRE(N >= 4 * kTileM && N % (4 * kTileM) == 0, L4P_ERR_SHAPE, "attention (CTA pair): N=%d must be a multiple of 512", N);
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)kDPad, (uint64_t)B * H * N};
    const uint64_t strides[1] = {(uint64_t)kDPad * 2};
    const uint32_t boxq[2] = {32, (uint32_t)kTileM};
    const uint32_t boxk[2] = {32, (uint32_t)kTileN / 2};
    rc = host_make_tmap_16b(&tmQ, q, 2, dims, strides, boxq, 64);
    if (rc != L4P_OK) return rc;
    rc = host_make_tmap_16b(&tmK, k, 2, dims, strides, boxk, 64);
    if (rc != L4P_OK) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)B * H * kDPad};
    const uint64_t strides[1] = {(uint64_t)N * 2};
    const uint32_t box[2] = {64, (uint32_t)kDPad / 2};
    rc = host_make_tmap_16b(&tmV, vt, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  }
  AttParams p;
  p.out = (uint16_t*)out;
  p.B = B; p.H = H; p.N = N; p.head_dim = head_dim;
  p.scale_log2 = scale * 1.4426950408889634f;
  typedef void (*KFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttParams);
  static const KFn table[2][2] = {{attention_pair_kernel<false, 0>, attention_pair_kernel<false, 2>},
                                  {attention_pair_kernel<true, 0>, attention_pair_kernel<true, 2>}};
  const int pi = poly == 0 ? 0 : 1;
  KFn kfn = table[bf16 ? 1 : 0][pi];
  static bool attr_set[2][2] = {};
  if (!attr_set[bf16 ? 1 : 0][pi]) {
    L4P_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    attr_set[bf16 ? 1 : 0][pi] = true;
  }
  const int grid = B * H * (N / (4 * kTileM)) * 2;
  L4P_CHECK_CUDA(launch_pdl(kfn, dim3(grid), dim3(kAttThreads), (size_t)kAttSmem, (cudaStream_t)stream, tmQ, tmK, tmV, p));
  return L4P_OK;
}

}  // namespace l4p
