// K4: fused spatio-temporal multi-head attention, softmax(Q K^T * scale) V, for sm_100a.
//
// Flash-style, never materialises the [N,N] score matrix (the reference does: modeling_finetune.py:180-186).
// One CTA owns two 128-row query tiles of one (batch, head) and streams the keys/values in blocks of 128:
//
//   warp 0       TMA producer: Q tiles once, then K blocks ([128 keys x 96] as three 64-byte-swizzled
//                32-element chunks) and V^T blocks ([96 x 128 keys], two 128-byte-swizzled chunks) in rings
//   warp 1       UMMA issuer (one thread): S_t = Q_t K_j^T (M128 N128 K96) into TMEM, O_t += P_t V_j
//                (M128 N96 K128) into TMEM; S_t(j+1) is issued before P_t(j) V_j so the tensor pipe
//                works while the softmax warps run
//   warp 2       TMEM allocator (512 columns: S0 | S1 | O0 | O1 | P0)
//   SPLIT = false (384 threads): warps 4..7 softmax warpgroup for tile 0, warps 8..11 for tile 1, one query row per thread
//   SPLIT = true  (640 threads, opt-in L4P_ATT_SPLIT=1): FOUR softmax warpgroups, two per tile: warps 4..7 / 8..11 own key columns
//                [0,64) / [64,128) of tile 0's rows, warps 12..15 / 16..19 the same for tile 1; the halves exchange their
//                partial row maximum through shared memory (one store, one 256-thread named barrier, one load per key
//                block) so both use the same reference maximum. Measured in round 2 (profiles/attention_r2.md): the exp
//                phase does not get shorter with half the scores per thread (the MUFU / FMA pipes of the sub-partition are
//                shared by the two halves, which run in lockstep), the exchange adds ~200 cycles: 241 vs 218 us at B = 8.
//
// Softmax: fp32 scores from TMEM, running max with lazy rescale (O is only rescaled in TMEM when the row max
// grew by more than 2^8), exp2 with the scale folded into one FFMA, fp32 row sums, P rounded to the operand
// type; tile 0's P goes back to TMEM (64 spare columns) as the A operand of a TS-mode PV UMMA, tile 1's P to
// 128B-swizzled smem (SS mode).
//
// Layouts (produced by the QKV GEMM epilogue, see gemm.cu L4P_STORE_QKV):
//   Q, K : [B, H, N, dpad]   (dpad = 96, columns >= head_dim are zero)
//   Vt   : [B, H, dpad, N]
//   out  : [B*N, H*head_dim] row-major 16-bit (operand of the output projection GEMM)
#include <cstdlib>

#include "common.cuh"

namespace l4p {

constexpr int kAttThreads = 384;    // SPLIT = false
constexpr int kAttThreads4 = 640;   // SPLIT = true
constexpr int kDPad = 96;
constexpr int kTileM = 128;   // query rows per tile
constexpr int kTileN = 128;   // keys per block
constexpr int kQTileBytes = kTileM * kDPad * 2;  // 24576: 3 chunks x (128 rows x 64 B)
constexpr int kKBytes = kTileN * kDPad * 2;      // 24576
constexpr int kVBytes = kDPad * kTileN * 2;      // 24576: 2 chunks x (96 rows x 128 B)
constexpr int kPBytes = kTileM * kTileN * 2;     // 32768: 2 chunks x (128 rows x 128 B)
constexpr int kKS = 2, kVS = 2;
constexpr int kAttSmem = 2 * kQTileBytes + kKS * kKBytes + kVS * kVBytes + 2 * kPBytes + 1024;
constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 352, kColP0 = 448;  // P0: tile 0's probabilities (64 columns)
constexpr float kRescaleThreshold = 8.0f;  // log2 units
// Variants measured on hardware in round 2 and removed (profiles/attention_r2.md): both tiles' P aliased onto S in TMEM
// (FA4-style; S(j+1) then queues behind PV(j): -11 %), a cta_group::2 kernel sharing K/V between two SMs (-20 %: the
// shared-memory operand port is NOT the limiter), S(j+1) of both tiles issued before PV(j) (-2 %), and the softmax
// denominator taken from a ones-row of V^T through the PV UMMA (+-0 %), an event-driven issuer that polls the four
// operand-ready conditions instead of the fixed S0 PV0 S1 PV1 order (-3 %).
struct AttParams {
  uint16_t* out;
  int B, H, N, head_dim;
  float scale_log2;  // scale * log2(e)
  long long* prof;   // optional [3 roles][64 iters][8] clock64 stamps of CTA 0 (debug; NULL in production)
};

#define ATT_STAMP(role, it, slot) \
  do { if (p.prof != nullptr && blockIdx.x == 0 && lane == 0 && (it) < 64) p.prof[((role) * 64 + (it)) * 8 + (slot)] = clock64(); } while (0)

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

template <bool BF16, int POLY, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? kAttThreads4 : kAttThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q;
  __shared__ __align__(8) uint64_t bar_kfull[kKS], bar_kempty[kKS];
  __shared__ __align__(8) uint64_t bar_vfull[kVS], bar_vempty[kVS];
  __shared__ __align__(8) uint64_t bar_sfull[2], bar_sfree[2], bar_pfull[2], bar_pvdone[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_xchg[SPLIT ? 2 : 1][2][2][kTileM];  // SPLIT: [key-block parity][tile][half][row] partial row maxima / sums
  constexpr uint32_t kSoftmaxThreads = SPLIT ? 256u : 128u;  // per tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.prof != nullptr && threadIdx.x == 0) {  // per-CTA wall/cycle stamps after the [3][64][8] timeline
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long* c = p.prof + 1536 + 5 * (long long)blockIdx.x;
    c[0] = gt; c[2] = clock64(); c[4] = smid;
  }
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + 2 * kQTileBytes;
  const uint32_t sV = sK + kKS * kKBytes;
  const uint32_t sP = sV + kVS * kVBytes;

  const int pairs_per_head = p.N / (2 * kTileM);
  const int pair = blockIdx.x % pairs_per_head;
  const int bh = blockIdx.x / pairs_per_head;  // b * H + h
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
      mbar_init(smem_u32(&bar_sfree[t]), kSoftmaxThreads);
      mbar_init(smem_u32(&bar_pfull[t]), kSoftmaxThreads);
      mbar_init(smem_u32(&bar_pvdone[t]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();  // Q/K/V are the previous kernel's outputs

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      const int q_row0 = bh * p.N + pair * 2 * kTileM;  // row in the [B*H*N, 96] view
      const uint32_t qb = smem_u32(&bar_q);
      mbar_expect_tx(qb, 2 * kQTileBytes);
      for (int t = 0; t < 2; ++t)
        for (int c = 0; c < 3; ++c)
          tma_load_2d(sQ + t * kQTileBytes + c * (kTileM * 64), &tmQ, qb, c * 32, q_row0 + t * kTileM);
      for (int j = 0; j < nblk; ++j) {
        {
          const int s = j % kKS;
          mbar_wait(smem_u32(&bar_kempty[s]), (((uint32_t)(j / kKS)) & 1u) ^ 1u);
          const uint32_t fb = smem_u32(&bar_kfull[s]);
          mbar_expect_tx(fb, kKBytes);
          for (int c = 0; c < 3; ++c)
            tma_load_2d(sK + s * kKBytes + c * (kTileN * 64), &tmK, fb, c * 32, bh * p.N + j * kTileN);
        }
        {
          const int s = j % kVS;
          mbar_wait(smem_u32(&bar_vempty[s]), (((uint32_t)(j / kVS)) & 1u) ^ 1u);
          const uint32_t fb = smem_u32(&bar_vfull[s]);
          mbar_expect_tx(fb, kVBytes);
          for (int c = 0; c < 2; ++c)
            tma_load_2d(sV + s * kVBytes + c * (kDPad * 128), &tmV, fb, j * kTileN + c * 64, bh * kDPad);
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- UMMA issuer
      // The whole warp walks the pipeline (warp-uniform control flow keeps descriptors in uniform registers);
      // one elected lane issues. Descriptor words are precomputed: per UMMA only an add remains.
      // Measured (tools/ubench/umma_bench.cu): an SS-mode UMMA costs ~43 + N/2 cycles (the 128 x 16 A operand is
      // fetched from shared memory before B streams), a TS-mode one ~10 + N/2, and one thread issues at most one per
      // ~93 cycles. The 12 + 16 UMMAs of a key block therefore occupy the tensor pipe for ~2700 cycles against 1536
      // cycles of arithmetic - that, not the softmax, bounds this kernel. A second issuer warp (one per query tile)
      // was tried: the pipe speeds up ~10 % but both softmax warpgroups then run their exp phases in lockstep and the
      // kernel gets slower.
      const bool leader = elect_one();
      const uint32_t idesc_s = umma_idesc_f16(BF16, kTileM, kTileN);
      const uint32_t idesc_o = umma_idesc_f16(BF16, kTileM, kDPad);
      constexpr uint32_t hi64 = umma_desc_hi(64, 4), hi128 = umma_desc_hi(128, 2);
      const uint32_t q_lo = umma_desc_lo(sQ), k_lo = umma_desc_lo(sK), p_lo = umma_desc_lo(sP), v_lo = umma_desc_lo(sV);
      auto issue_s = [&](const int t, const int s) {
        const uint32_t d = tmem_base + (t == 0 ? kColS0 : kColS1);
        const uint32_t qa = q_lo + (uint32_t)t * (kQTileBytes >> 4), ka = k_lo + (uint32_t)s * (kKBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < kDPad / 16; ++kk) {
          const uint32_t off = ((uint32_t)(kk >> 1) * (kTileM * 64) + (uint32_t)(kk & 1) * 32) >> 4;
          umma_ss(d, umma_desc_make(qa + off, hi64), umma_desc_make(ka + off, hi64), idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar_sfull[t]));
      };
      auto issue_pv = [&](const int t, const int s, const uint32_t acc) {
        const uint32_t d = tmem_base + (t == 0 ? kColO0 : kColO1);
        const uint32_t pa = p_lo + (uint32_t)t * (kPBytes >> 4), va = v_lo + (uint32_t)s * (kVBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < kTileN / 16; ++kk) {
          const uint32_t o = ((uint32_t)(kk & 3) * 32) >> 4;
          const uint64_t vdesc = umma_desc_make(va + (uint32_t)(kk >> 2) * ((kDPad * 128) >> 4) + o, hi128);
          if (t == 0)  // P_0 is the TMEM A operand: 8 columns per K = 16 step, no shared-memory traffic for P
            umma_ts(d, tmem_base + kColP0 + (uint32_t)kk * 8u, vdesc, idesc_o, kk != 0 ? 1u : acc);
          else
            umma_ss(d, umma_desc_make(pa + (uint32_t)(kk >> 2) * ((kTileM * 128) >> 4) + o, hi128), vdesc, idesc_o,
                    kk != 0 ? 1u : acc);
        }
        umma_commit(smem_u32(&bar_pvdone[t]));
      };

      mbar_wait(smem_u32(&bar_q), 0);
      mbar_wait(smem_u32(&bar_kfull[0]), 0);
      tc_fence_after();
      if (leader) {
        issue_s(0, 0);
        issue_s(1, 0);
        umma_commit(smem_u32(&bar_kempty[0]));
      }
      __syncwarp();
      for (int j = 0; j < nblk; ++j) {
        const int jn = j + 1;
        const int sk = jn % kKS, sv = j % kVS;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (jn < nblk) {
            if (t == 0) mbar_wait(smem_u32(&bar_kfull[sk]), ((uint32_t)(jn / kKS)) & 1u);
            mbar_wait(smem_u32(&bar_sfree[t]), (uint32_t)j & 1u);  // softmax t holds S_t(j) in registers
            tc_fence_after();
            if (leader) {
              issue_s(t, sk);
              if (t == 1) umma_commit(smem_u32(&bar_kempty[sk]));
            }
            __syncwarp();
          }
          ATT_STAMP(2, j, t * 4 + 0);
          mbar_wait(smem_u32(&bar_pfull[t]), (uint32_t)j & 1u);  // P_t(j) ready (and O_t rescaled)
          ATT_STAMP(2, j, t * 4 + 1);
          if (t == 0) mbar_wait(smem_u32(&bar_vfull[sv]), ((uint32_t)(j / kVS)) & 1u);
          tc_fence_after();
          if (leader) {
            issue_pv(t, sv, j != 0 ? 1u : 0u);
            if (t == 1) umma_commit(smem_u32(&bar_vempty[sv]));
          }
          __syncwarp();
          ATT_STAMP(2, j, t * 4 + 2);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    // NC = key columns of a 128-key block owned by one thread: the whole row (SPLIT = false) or half of it (SPLIT = true)
    constexpr int NC = SPLIT ? 64 : 128;
    // setmaxnreg moves registers inside the CTA's launch-time allocation only (640 threads x 96): 128 x 56 + 512 x 104 fits,
    // 112 does not (the third warpgroup's request would block forever)
    if (SPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int wg = (warp - 4) >> 2;            // softmax warpgroup
    const int t = SPLIT ? (wg >> 1) : wg;      // tile
    const int h = SPLIT ? (wg & 1) : 0;        // column half
    const int q4 = warp & 3;                   // TMEM lane quarter
    const int r = q4 * 32 + lane;              // row in tile
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + (t == 0 ? kColS0 : kColS1) + (uint32_t)(h * NC);
    const uint32_t tO = tmem_base + lane_addr + (t == 0 ? kColO0 : kColO1);
    const uint32_t pRow = sP + t * kPBytes + (uint32_t)r * 128u;
    const uint32_t swz = (uint32_t)(r & 7);
    const float c = p.scale_log2;
    const bool stamp = h == 0;                 // timeline role = tile
    // this thread's share of the O row (rescale + epilogue): all 96 columns, or 48 per half
    constexpr int OC = SPLIT ? kDPad / 2 : kDPad;
    const uint32_t tOmine = tO + (uint32_t)(h * OC);

    float m_used = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < nblk; ++j) {
      if (stamp) ATT_STAMP(t, j, 0);
      mbar_wait(smem_u32(&bar_sfull[t]), (uint32_t)j & 1u);
      tc_fence_after();
      if (stamp) ATT_STAMP(t, j, 1);
      uint32_t s[NC];
#pragma unroll
      for (int cc = 0; cc < NC; cc += 32) tmem_ld32(tS + cc, s + cc);
      tmem_ld_wait();
      if (stamp) ATT_STAMP(t, j, 2);
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_sfree[t]));

      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]),
            mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i < NC; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
      }
      float mxr = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if (SPLIT) {
        // both halves of a row must use the same reference maximum (they feed the same O accumulator): exchange the partial
        // maxima. Double-buffered by key-block parity: the partner reads buffer (j & 1) after barrier j, this thread's next
        // write to the same buffer happens after barrier j + 1, which the partner only reaches after that read.
        s_xchg[j & 1][t][h][r] = mxr;
        named_bar_sync(1u + (uint32_t)t, kSoftmaxThreads);
        mxr = fmaxf(mxr, s_xchg[j & 1][t][h ^ 1][r]);
      }
      const float mx = mxr * c;

      if (j == 0) {
        m_used = mx;
      } else {
        const bool need = mx > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // rare: rescale the running output in TMEM (whole warp, each row with its own factor; the partner warp of the
          // other half sees the same 32 row maxima, takes the same branch and rescales its own 48 columns)
          const float m_new = fmaxf(m_used, mx);
          const float alpha = ex2(m_used - m_new);
          mbar_wait(smem_u32(&bar_pvdone[t]), (uint32_t)(j - 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int cc = 0; cc < OC; cc += 16) {
            uint32_t o[16];
            tmem_ld16(tOmine + cc, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tOmine + cc, o);
          }
          tmem_st_wait();
          tc_fence_before();
          l *= alpha;
          m_used = m_new;
        }
      }

      if (stamp) ATT_STAMP(t, j, 3);
      // p = exp2(s*c - m_used), row sum in fp32, pack pairs in place
      // scale/subtract and the row sum run as packed f32x2; of every 16 scores, 2*POLY take the polynomial exp2 on the
      // FMA pipe and the rest the MUFU, which balances the two pipes (16 MUFU results per clock and SM would otherwise need
      // 2048 cycles per key block, more than the ~1600 cycles of its UMMAs)
      const uint64_t c2 = pk2(c, c), nm2 = pk2(-m_used, -m_used);
      uint64_t lsum0 = pk2(0.f, 0.f), lsum1 = pk2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < NC; i += 2) {
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
      if (stamp) ATT_STAMP(t, j, 4);

      if (j > 0) mbar_wait(smem_u32(&bar_pvdone[t]), (uint32_t)(j - 1) & 1u);  // P_t is free again
      if (stamp) ATT_STAMP(t, j, 5);
      if (t == 0) {
        // tile 0: P goes straight back to TMEM (row = lane, two probabilities per 32-bit column) as the A operand of PV
        const uint32_t tP = tmem_base + lane_addr + kColP0 + (uint32_t)(h * (NC / 2));
#pragma unroll
        for (int cc = 0; cc < NC / 2; cc += 16) tmem_st16(tP + cc, s + cc);
        tmem_st_wait();
        tc_fence_before();
      } else {
#pragma unroll
        for (int u = 0; u < NC / 8; ++u) {
          // 16-byte unit: keys [8 uu, 8 uu + 8) of the block; chunk = uu / 8 (= the column half when SPLIT); swizzled unit
          // inside the 128-byte row
          const int uu = u + h * (NC / 8);
          const uint32_t addr = pRow + (uint32_t)(uu >> 3) * (kTileM * 128) + ((((uint32_t)uu & 7u) ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(s[4 * u]), "r"(s[4 * u + 1]),
                       "r"(s[4 * u + 2]), "r"(s[4 * u + 3])
                       : "memory");
        }
        fence_proxy_async();
      }
      mbar_arrive(smem_u32(&bar_pfull[t]));
      if (stamp) ATT_STAMP(t, j, 6);
    }

    // ---- epilogue: O_t / l -> global
    mbar_wait(smem_u32(&bar_pvdone[t]), (uint32_t)(nblk - 1) & 1u);
    tc_fence_after();
    if (SPLIT) {  // row sum = sum of both halves (buffer of parity nblk: not in use by the last key block's exchange)
      s_xchg[nblk & 1][t][h][r] = l;
      named_bar_sync(1u + (uint32_t)t, kSoftmaxThreads);
      l += s_xchg[nblk & 1][t][h ^ 1][r];
    }
    const float inv_l = 1.0f / l;
    const int b = bh / p.H, hh = bh - b * p.H;
    const long long row = (long long)b * p.N + (long long)pair * 2 * kTileM + t * kTileM + r;
    uint16_t* dst = p.out + row * ((long long)p.H * p.head_dim) + (long long)hh * p.head_dim;
#pragma unroll
    for (int cc = 0; cc < OC; cc += 16) {
      uint32_t o[16];
      tmem_ld16(tOmine + cc, o);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = h * OC + cc + g * 8;
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

  __syncthreads();
  if (p.prof != nullptr && threadIdx.x == 0) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    long long* c = p.prof + 1536 + 5 * (long long)blockIdx.x;
    c[1] = gt; c[3] = clock64();
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace l4p

using namespace l4p;

extern "C" int l4p_attention(const void* q, const void* k, const void* vt, void* out, int B, int H, int N,
                             int head_dim, int head_dim_pad, float scale, int bf16, void* stream, void* prof) {
  L4P_REQUIRE(q && k && vt && out, L4P_ERR_ARG, "l4p_attention: null pointer");
  L4P_REQUIRE(B > 0 && H > 0, L4P_ERR_SHAPE, "l4p_attention: B=%d H=%d", B, H);
  L4P_REQUIRE(head_dim_pad == kDPad && head_dim % 8 == 0 && head_dim > 0 && head_dim <= kDPad, L4P_ERR_SHAPE,
              "l4p_attention: head_dim=%d pad=%d (this build: pad 96, head_dim multiple of 8 <= 96)", head_dim,
              head_dim_pad);
  L4P_REQUIRE(N >= 2 * kTileM && N % (2 * kTileM) == 0, L4P_ERR_SHAPE, "l4p_attention: N=%d must be a multiple of 256", N);
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)kDPad, (uint64_t)B * H * N};
    const uint64_t strides[1] = {(uint64_t)kDPad * 2};
    const uint32_t box[2] = {32, (uint32_t)kTileM};
    rc = host_make_tmap_16b(&tmQ, q, 2, dims, strides, box, 64);
    if (rc != L4P_OK) return rc;
    rc = host_make_tmap_16b(&tmK, k, 2, dims, strides, box, 64);
    if (rc != L4P_OK) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)B * H * kDPad};
    const uint64_t strides[1] = {(uint64_t)N * 2};
    const uint32_t box[2] = {64, (uint32_t)kDPad};
    rc = host_make_tmap_16b(&tmV, vt, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  }
  AttParams p;
  p.out = (uint16_t*)out;
  p.B = B; p.H = H; p.N = N; p.head_dim = head_dim;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.prof = (long long*)prof;
  // POLY = number of score pairs (of every 8) whose exp2 runs as an FMA-pipe polynomial instead of MUFU
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("L4P_ATT_POLY");
    poly = e ? atoi(e) : 2;
    if (poly < 0 || poly > 3) poly = 0;
  }
  // 2 full-row softmax warpgroups by default; L4P_ATT_SPLIT=1 selects the 4 half-row warpgroup variant (measured 10 % slower,
  // kept parity-tested: tests/test_gemm_gpu.py::test_attention_split_variant)
  static int split = -1;
  if (split < 0) {
    const char* e = getenv("L4P_ATT_SPLIT");
    split = (e && atoi(e) == 1) ? 1 : 0;
  }
  typedef void (*KFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttParams);
  static const KFn table[2][2][4] = {
      {{attention_kernel<false, 0, false>, attention_kernel<false, 1, false>, attention_kernel<false, 2, false>,
        attention_kernel<false, 3, false>},
       {attention_kernel<true, 0, false>, attention_kernel<true, 1, false>, attention_kernel<true, 2, false>,
        attention_kernel<true, 3, false>}},
      {{attention_kernel<false, 0, true>, attention_kernel<false, 1, true>, attention_kernel<false, 2, true>,
        attention_kernel<false, 3, true>},
       {attention_kernel<true, 0, true>, attention_kernel<true, 1, true>, attention_kernel<true, 2, true>,
        attention_kernel<true, 3, true>}}};
  KFn kfn = table[split][bf16 ? 1 : 0][poly];
  static bool attr_set[2][2][4] = {};
  if (!attr_set[split][bf16 ? 1 : 0][poly]) {
    L4P_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    attr_set[split][bf16 ? 1 : 0][poly] = true;
  }
  const int grid = B * H * (N / (2 * kTileM));
  L4P_CHECK_CUDA(launch_pdl(kfn, dim3(grid), dim3(split ? kAttThreads4 : kAttThreads), (size_t)kAttSmem, (cudaStream_t)stream, tmQ, tmK, tmV, p));
  return L4P_OK;
}
