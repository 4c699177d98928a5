// Blackwell (sm_100a) device primitives shared by every kernel in this library:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA + TMEM) and small numeric helpers.
// Everything here is inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/l4p_b200.h"

namespace l4p {

#define L4P_DEVICE __device__ __forceinline__

// ----------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------
L4P_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

L4P_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 %%rx;\n\t"
      ".reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %1;\n\t"
      "@%%px mov.s32 %0, 1;\n\t"
      "}\n"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its stream predecessor is
// still running; pdl_wait() blocks until the predecessor grid has completed and its memory is visible. Everything
// before pdl_wait() (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail and hides
// the ~3 us launch latency. pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as SMs free up.
// ----------------------------------------------------------------------------------------
L4P_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
L4P_DEVICE void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
L4P_DEVICE void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
L4P_DEVICE void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
L4P_DEVICE void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
L4P_DEVICE void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
L4P_DEVICE bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error on the host) instead of hanging the GPU.
#ifndef L4P_MBAR_TIMEOUT_CYCLES
#define L4P_MBAR_TIMEOUT_CYCLES (4000000000ll)
#endif
L4P_DEVICE void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > L4P_MBAR_TIMEOUT_CYCLES) {
      printf("l4p: mbarrier timeout block(%d,%d,%d) thread %d bar 0x%x parity %u\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA)
L4P_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// named barrier among a subset of warps
L4P_DEVICE void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------
L4P_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
L4P_DEVICE void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
L4P_DEVICE void tma_load_3d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
L4P_DEVICE void tma_load_5d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, UMMA, commit, ld/st, fences
// ----------------------------------------------------------------------------------------
L4P_DEVICE void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
L4P_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
L4P_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
L4P_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; single thread issues.
L4P_DEVICE void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M x 16 16-bit elements per instruction) sits in TMEM, one row per
// lane, two elements per 32-bit column (8 columns per K = 16 step)
L4P_DEVICE void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued UMMAs of this thread have completed
L4P_DEVICE void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// K-major shared-memory operand descriptor. Rows are `row_bytes` wide (= the swizzle span: 128/64/32 B),
// 8-row groups are `8*row_bytes` apart (SBO). layout_type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B.
L4P_DEVICE uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t row_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // start address  [0,14)
  d |= (uint64_t)1 << 16;                                 // LBO (ignored for swizzled K-major) [16,30)
  d |= (uint64_t)(((8u * row_bytes) >> 4) & 0x3FFF) << 32;  // SBO [32,46)
  d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell) [46,48)
  d |= (uint64_t)(layout_type & 7) << 61;                 // layout type [61,64)
  return d;
}
// Split form for hot issue loops: the high word is constant per layout, the low word is (addr >> 4) | LBO.
L4P_DEVICE constexpr uint32_t umma_desc_hi(uint32_t row_bytes, uint32_t layout_type) {
  return (((8u * row_bytes) >> 4) & 0x3FFFu) | (1u << 14) | ((layout_type & 7u) << 29);
}
L4P_DEVICE uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
L4P_DEVICE uint64_t umma_desc_make(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// Instruction descriptor for kind::f16 (fp16/bf16 operands, fp32 accumulate), both operands K-major.
L4P_DEVICE uint32_t umma_idesc_f16(bool bf16, uint32_t M, uint32_t N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= (bf16 ? 1u : 0u) << 7;      // a_format
  d |= (bf16 ? 1u : 0u) << 10;     // b_format
  d |= (N >> 3) << 17;             // n_dim
  d |= (M >> 4) << 24;             // m_dim
  return d;
}

// TMEM -> registers: this warp's 32 lanes x N consecutive 32-bit columns (thread i <-> lane base+i)
L4P_DEVICE void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr));
}
L4P_DEVICE void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
L4P_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
L4P_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
L4P_DEVICE void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
      "r"(r[15])
      : "memory");
}

// ----------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): cluster helpers, 2-SM TMA / UMMA / commit / TMEM allocation
// ----------------------------------------------------------------------------------------
L4P_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
L4P_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem location in CTA `rank` of this cluster
L4P_DEVICE uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
L4P_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on an mbarrier that may live in the peer CTA of the pair
L4P_DEVICE void tma2_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
L4P_DEVICE void tma2_load_3d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
L4P_DEVICE void tma2_load_5d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2, int c3,
                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
L4P_DEVICE void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair, same smem_dst offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
L4P_DEVICE void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]; issued by the leader CTA only
L4P_DEVICE void umma2_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior UMMAs of this thread completed) on the barrier at this smem offset in every CTA of `mask`
L4P_DEVICE void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// ----------------------------------------------------------------------------------------
// numerics
// ----------------------------------------------------------------------------------------
template <bool BF16>
L4P_DEVICE uint32_t pack2(float a, float b) {  // a -> low 16 bits (lower address)
  if constexpr (BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
template <bool BF16>
L4P_DEVICE float2 unpack2(uint32_t u) {
  if constexpr (BF16) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
}
template <bool BF16>
L4P_DEVICE uint16_t pack1(float a) {
  if constexpr (BF16) {
    __nv_bfloat16 v = __float2bfloat16_rn(a);
    return *reinterpret_cast<uint16_t*>(&v);
  } else {
    __half v = __float2half_rn(a);
    return *reinterpret_cast<uint16_t*>(&v);
  }
}
template <bool BF16>
L4P_DEVICE float unpack1(uint16_t u) {
  if constexpr (BF16) {
    return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
  } else {
    return __half2float(*reinterpret_cast<__half*>(&u));
  }
}
L4P_DEVICE float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7 + fast-intrinsic error ~1e-6): ~14 instructions instead of
// erff's ~40. Used where the result is rounded to a 16-bit operand anyway (2^-11 relative).
L4P_DEVICE float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float y = 1.0f - p * t * __expf(-ax * ax);
  return copysignf(y, x);
}
L4P_DEVICE float gelu_erf_fast(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f)); }

L4P_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
L4P_DEVICE float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32x2 math (Blackwell FFMA2/FADD2): two elements per FMA-pipe issue slot
L4P_DEVICE uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
L4P_DEVICE void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
L4P_DEVICE uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
L4P_DEVICE uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
L4P_DEVICE uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// exact (erf) GELU of two values, arithmetic packed two-wide, ONE MUFU per value:
//   gelu(x) = x Phi(x) = max(x, 0) - |x| * 0.5 * erfc(|x| / sqrt 2),   erfc(|x| / sqrt 2) = 2^r(|x|),  r(a) = a q(a)
// with q a degree-7 polynomial fitted (weighted minimax, fp32 Horner evaluation included in the fit) to log2 erfc on
// [0, 6.5]: |gelu error| <= 4.2e-8 absolute for all x; beyond 6.5 r keeps falling (r <= -33.8, finite up to the fp16
// maximum), so the tail term vanishes by itself and no clamp is needed. The previous form (A&S 7.1.26: rcp + exp) needed two
// MUFU operations per value, which made the GELU epilogues MUFU-bound (16 results / clk / SM): 2816 cycles per 128 x 176 tile
// of the mask decoder's hyper ConvT against 2100 cycles of tensor work.
L4P_DEVICE void gelu2(float& x0, float& x1) {
  const uint64_t ax = pk2(fabsf(x0), fabsf(x1));
  uint64_t q = fma2(pk2(-1.9019885257876012e-06f, -1.9019885257876012e-06f), ax, pk2(2.805595431709662e-05f, 2.805595431709662e-05f));
  q = fma2(q, ax, pk2(-0.00013146817218512297f, -0.00013146817218512297f));
  q = fma2(q, ax, pk2(-0.00027208906249143183f, -0.00027208906249143183f));
  q = fma2(q, ax, pk2(0.007245440501719713f, 0.007245440501719713f));
  q = fma2(q, ax, pk2(-0.052627626806497574f, -0.052627626806497574f));
  q = fma2(q, ax, pk2(-0.4591621458530426f, -0.4591621458530426f));
  q = fma2(q, ax, pk2(-1.1511110067367554f, -1.1511110067367554f));
  float r0, r1;
  upk2(mul2(q, ax), r0, r1);
  const uint64_t e = pk2(ex2(r0), ex2(r1));                       // erfc(|x| / sqrt 2)
  const uint64_t m = pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
  upk2(fma2(mul2(ax, pk2(-0.5f, -0.5f)), e, m), x0, x1);
}

L4P_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
L4P_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------------------
// host helpers
// ----------------------------------------------------------------------------------------
int host_set_error(int code, const char* fmt, ...);   // records l4p_last_error(), returns code
int host_check_cuda(cudaError_t e, const char* what); // L4P_OK or L4P_ERR_CUDA

// Encode a tiled TMA descriptor for a 16-bit element tensor. dims/strides innermost first;
// strides_bytes has rank-1 entries (stride of dim 1.. in bytes). swizzle_bytes in {0,32,64,128}.
int host_make_tmap_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);
int host_num_sms();
bool host_pdl_enabled();  // false when L4P_NO_PDL=1

// Launch with the programmatic-stream-serialization attribute (see pdl_wait above). The kernel MUST call pdl_wait()
// before its first global-memory access.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = host_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

#define L4P_CHECK_CUDA(expr)                                  \
  do {                                                        \
    int _rc = ::l4p::host_check_cuda((expr), #expr);          \
    if (_rc != L4P_OK) return _rc;                     \
  } while (0)
#define L4P_REQUIRE(cond, code, ...)                                \
  do {                                                              \
    if (!(cond)) return ::l4p::host_set_error((code), __VA_ARGS__); \
  } while (0)

}  // namespace l4p
