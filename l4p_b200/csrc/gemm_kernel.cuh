// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3-D convolution kernels for sm_100a (device side).
//
//   D[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
//   warps 0..7  epilogue       (tcgen05.ld -> bias/activation/residual -> global stores); two warpgroups interleave
//               column chunks so that the latency-bound epilogue of one overlaps the other
//   warp 8      TMA producer   (A/B tiles -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 9      UMMA issuer    (tcgen05.mma kind::f16, fp32 accumulators in TMEM, 2 accumulator stages)
//   warp 10     TMEM allocator
// The two single-thread pipeline warps sit at the HIGHEST warp indices: the SM's issue arbiter favours high warp ids,
// and a producer / MMA issuer starved by eight busy epilogue warps stalls the tensor pipe.
//
// A-operand modes:
//   A_MATRIX  plain row-major [M,K] matrix, one 2-D TMA box (64 x 128) per k-block.
//   A_CONV3D  channels-last activations [B,T,H,W,C]; a 128-row tile is a (bT,bH,bW) voxel box and the k-loop
//             runs over (filter tap, 64-channel block): every tap is the same 5-D TMA box shifted by the tap
//             offset, the zero padding comes from TMA out-of-bounds fill. No im2col buffer exists.
//
// The epilogue is specialised at COMPILE TIME (template parameter EPI, see epi_* below): a single warp executes it
// row by row with little latency hiding, so every runtime mode test / parameter reload inside its loops costs tens of
// cycles. gemm.cu maps a descriptor to the matching instance; EPI_GENERIC instances keep all ROWMAJOR options runtime.
//
// Reference call sites replaced: modeling_finetune.py:62-69,171-177,188 (Linear), dpt_block.py:29-90,
// 144-157,255-278,406-414 (Conv3d / ConvTranspose3d), sam/transformer.py:223-245.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace l4p {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // 64 x 2 B = one 128 B swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kAccCols = 256;                     // TMEM columns per accumulator stage
constexpr int kMaxEpiGroups = 4;
constexpr int kEpiChunk = 32;                     // columns per transposed chunk
constexpr int kEpiStageBytes = 32 * 128;          // per-warp staging: 32 rows x 32 fp32
constexpr int epi_smem_bytes(int groups) { return 4 * groups * kEpiStageBytes; }

// ---- compile-time epilogue configuration -------------------------------------------------------------------------
//   bits 0-2 store mode (L4P_STORE_*), bits 3-4 activation (L4P_ACT_*), then ROWMAJOR option flags
constexpr int EPI_RES32 = 1 << 5;    // fp32 residual (optionally a broadcast table: res_row_mod)
constexpr int EPI_RES16 = 1 << 6;    // one or two 16-bit residuals (null-checked at run time)
constexpr int EPI_OUT32 = 1 << 7;
constexpr int EPI_OUT16 = 1 << 8;
constexpr int EPI_OUT16R = 1 << 9;   // second 16-bit output = relu(out)
constexpr int EPI_GENERIC = 1 << 10; // activation and ROWMAJOR options decided at run time (slow, always correct)
constexpr int EPI_WIDE3 = 1 << 11;   // THREE epilogue warpgroups (512 threads, 128 registers each): for short-K problems, whose time is
                                     // the epilogue's (tools/outk48_prof.py: ~6 k cycles per 128 x 176 tile with two warpgroups)
constexpr int kStoreSplitK = 5;      // internal store mode: fp32 atomic accumulation of a K-slice into the split-K workspace
constexpr int epi_make(int store, int act, int flags) { return store | (act << 3) | flags; }
constexpr int epi_store(int e) { return e & 7; }
// Epilogue warpgroups (group g handles column chunks c with c % groups == g). The fused-dot modes are bound by the
// instruction throughput of their epilogue (GELU + per-row dot products) and need few registers: four groups (640
// threads); the store modes keep two groups with 200 registers each for the transposed sub-tile and its residuals.
constexpr int epi_groups(int e) {
  return (epi_store(e) == L4P_STORE_HEAD1X1 || epi_store(e) == L4P_STORE_HYPER) ? 4 : ((e & EPI_WIDE3) ? 3 : 2);
}
constexpr int gemm_threads(int e) { return 128 + 128 * epi_groups(e); }
constexpr int epi_act(int e) { return (e >> 3) & 3; }

// n / d for 0 <= n < 2^31 by multiply-high + shift (host-initialised, CUTLASS FastDivmod construction)
struct FastDiv {
  uint32_t mul, shr;
  int d;
#ifdef __CUDACC__
  L4P_DEVICE uint32_t div(uint32_t n) const { return d != 1 ? (__umulhi(n, mul) >> shr) : n; }
  L4P_DEVICE void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * (uint32_t)d; }
#endif
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  f.mul = 0; f.shr = 0;
  if (d > 1) {
    int lg = 0;
    while ((1ll << lg) < d) ++lg;
    const int pw = 31 + lg;
    f.mul = (uint32_t)(((1ull << pw) + (unsigned long long)d - 1) / (unsigned long long)d);
    f.shr = (uint32_t)(pw - 32);
  }
  return f;
}

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
  int ld_res;
  int res_row_mod;
  int store_mode;
  float* out_f32;
  uint16_t* out_16;
  uint16_t* out_16_relu;
  int ld_out;
  uint16_t *q, *k, *vt;
  int heads, head_dim, head_dim_pad, tokens;
  int sT, sH, sW, ctCout;
  const float* w2;
  const float* b2;
  int c2, exp_out;
  long long rows_per_group;  // STORE_HYPER: w2 is indexed by row / rows_per_group
  FastDiv fd_cW, fd_cH, fd_cT, fd_rpg;  // HYPER finalisation: row -> (g,t,h,w), row -> hyper group
  FastDiv fd_bW, fd_bH;                 // HEAD1X1 finalisation: row in tile -> (tl,hl,wl)
  int split_k;               // >1: work unit = (tile, k-range); partial sums are atomically added to splitk_ws [M,N] fp32
  float* splitk_ws;
  int k128;                  // matrix mode: one ring stage holds TWO 64-wide k-blocks per operand (one 3-D TMA box each, slab-major): halves the
                             // per-k-block cost of the single-thread producer / issuer loops (tools/pair_n_sweep.py)
  int a_halo;                // 2-CTA conv mode, 3x3 in-plane filter: one A box with bH+2 lines per (dt, dw, channel block) serves the 3 dh taps
  int conv_grp_b;            // conv mode, grouped weights: batch entries per group (0 = one shared W); group = b / conv_grp_b owns W rows
                             // [group * grp_b_rows, ...) and bias [group * N, ...): several heads' identical layers in one launch
  int m_stride;              // matrix mode: rows between consecutive M tiles = rows stored per tile (128 unless grouped)
  int grp_a_rows, grp_b_rows;  // grouped weights: tile rows / grp_a_rows = group, its W block starts at row group * grp_b_rows
  long long* prof;           // optional [3][512] clock64 timeline of CTA 0
};

#define GEMM_STAMP(role, idx) \
  do { if (p.prof != nullptr && blockIdx.x == 0 && (idx) < 512) p.prof[(role) * 512 + (idx)] = clock64(); } while (0)
// fine-grained epilogue stamps of thread 0 (slots 64.. of role 2); compiled in only for tuning builds
#ifndef L4P_GEMM_FINE_PROF
#define L4P_GEMM_FINE_PROF 0
#endif
#if L4P_GEMM_FINE_PROF
// stamps are collected in shared memory (a global read-modify-write per stamp would cost ~1000 cycles and distort the
// timeline) and flushed by FINE_FLUSH() at kernel end: role 2 slots 64.. (epilogue thread 0), role 0 slots 300.. (helper)
__shared__ long long g_fine[2][200];
__shared__ int g_fine_n[2];
#define FINE_INIT() do { if (threadIdx.x == 0) { g_fine_n[0] = 0; g_fine_n[1] = 0; } } while (0)
#define EPI_STAMP()                                                                       \
  do {                                                                                    \
    if (threadIdx.x == 0 && blockIdx.x == 0 && p.prof != nullptr) {                       \
      const int n_ = g_fine_n[0];                                                         \
      if (n_ < 200) { g_fine[0][n_] = clock64(); g_fine_n[0] = n_ + 1; }                  \
    }                                                                                     \
  } while (0)
#define HELPER_STAMP()                                                                    \
  do {                                                                                    \
    if (lane == 0 && blockIdx.x == 0 && p.prof != nullptr) {                              \
      const int n_ = g_fine_n[1];                                                         \
      if (n_ < 200) { g_fine[1][n_] = clock64(); g_fine_n[1] = n_ + 1; }                  \
    }                                                                                     \
  } while (0)
#define FINE_FLUSH()                                                                      \
  do {                                                                                    \
    if (threadIdx.x == 0 && blockIdx.x == 0 && p.prof != nullptr) {                       \
      for (int i_ = 0; i_ < g_fine_n[0]; ++i_) p.prof[2 * 512 + 64 + i_] = g_fine[0][i_]; \
      for (int i_ = 0; i_ < g_fine_n[1]; ++i_) p.prof[0 * 512 + 300 + i_] = g_fine[1][i_]; \
    }                                                                                     \
  } while (0)
#else
#define EPI_STAMP() do { } while (0)
#define HELPER_STAMP() do { } while (0)
#define FINE_INIT() do { } while (0)
#define FINE_FLUSH() do { } while (0)
#endif

// -DL4P_GEMM_GRID_PROF=1 (tools/gemm_grid_prof.py, experiment builds only): every CTA writes the global timer at kernel entry, after the
// PDL wait and at its end to prof[1536 + 3 * blockIdx.x ...] (a 2048-entry prof buffer): launch skew, slowest CTA, boundary gaps.
#ifdef L4P_GEMM_GRID_PROF
#define GRID_STAMP(i)                                                                          \
  do {                                                                                         \
    if (p.prof != nullptr && threadIdx.x == 0) {                                               \
      unsigned long long gt_;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                  \
      p.prof[1536 + 3 * blockIdx.x + (i)] = (long long)gt_;                                    \
    }                                                                                          \
  } while (0)
#else
#define GRID_STAMP(i) do { } while (0)
#endif

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

L4P_DEVICE void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
L4P_DEVICE float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// raw 8-byte global load of four 16-bit residual values (explicit state space: behind the null checks of the optional
// residual pointers the compiler otherwise falls back to generic LD). Not volatile: it may be scheduled freely, the values
// are only consumed one chunk later.
L4P_DEVICE uint2 ldg_res16x4(const uint16_t* p) {
  uint2 v;
  asm("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
// Predicated global accesses of the store-mode epilogue. As plain C++ inside `if (row / column valid)` every access became a branch
// with its own reconvergence barrier and re-derived its address from the kernel parameters (~35 SASS instructions per 4-row
// store, tools/epi_fine_prof.py); as predicated PTX there is no branch and the address is one IMAD.WIDE + LEA pair.
L4P_DEVICE void stg_f32x4_if(float* p, const float4& v, const bool ok) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q st.global.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)ok));
}
L4P_DEVICE void stg_b32x2_if(uint16_t* p, const uint32_t lo, const uint32_t hi, const bool ok) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.global.v2.b32 [%0], {%1, %2};\n\t}"
               ::"l"(p), "r"(lo), "r"(hi), "r"((int)ok));
}
L4P_DEVICE float4 ldg_f32x4_if(const float* p, const bool ok) {   // zeros when !ok; not volatile (consumed one chunk later)
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.global.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) : "l"(p), "r"((int)ok));
  return v;
}
L4P_DEVICE uint2 ldg_res16x4_if(const uint16_t* p, const bool ok) {   // +0.0 pairs when !ok
  uint2 v = make_uint2(0u, 0u);
  asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.global.v2.u32 {%0, %1}, [%2];\n\t}" : "+r"(v.x), "+r"(v.y) : "l"(p), "r"((int)ok));
  return v;
}
template <bool BF16>
L4P_DEVICE void add_res16(float4& a, const uint16_t* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 lo = unpack2<BF16>(u.x), hi = unpack2<BF16>(u.y);
  a.x += lo.x; a.y += lo.y; a.z += hi.x; a.w += hi.y;
}
template <bool BF16>
L4P_DEVICE void store4_16(uint16_t* p, const float4& v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack2<BF16>(v.x, v.y), pack2<BF16>(v.z, v.w));
}
// activation of a pair (the GELU runs two-wide on the packed f32x2 pipe)
template <int ACT>
L4P_DEVICE void apply_act2(float& a, float& b) {
  if constexpr (ACT == L4P_ACT_GELU) gelu2(a, b);
  else if constexpr (ACT == L4P_ACT_RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
}
L4P_DEVICE void apply_act2_rt(float& a, float& b, int act) {
  if (act == L4P_ACT_GELU) gelu2(a, b);
  else if (act == L4P_ACT_RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
}

// split-K: bias / activation / residuals / stores of four finished columns of one output row (run-time options: the
// split-K problems are the small low-resolution ones, their finalisation is latency-, not issue-bound)
template <bool BF16>
L4P_DEVICE void splitk_finalize4(const GemmKParams& p, const long long row, const int col, float4 x) {
  if (p.bias != nullptr) {
    int bg = 0;
    if (p.conv_grp_b > 0) {   // grouped conv: the row's batch entry selects the bias vector
      const long long vox = (long long)p.cT * p.cH * p.cW;
      bg = (int)((row / vox) / p.conv_grp_b) * p.N;
    }
    const float4 b4 = *reinterpret_cast<const float4*>(p.bias + bg + col);
    x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
  }
  apply_act2_rt(x.x, x.y, p.act);
  apply_act2_rt(x.z, x.w, p.act);
  if (p.res_f32 != nullptr) {
    const long long rrow = p.res_row_mod > 0 ? row % p.res_row_mod : row;
    const float4 r4 = *reinterpret_cast<const float4*>(p.res_f32 + rrow * p.ld_res + col);
    x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
  }
  if (p.res_16 != nullptr) add_res16<BF16>(x, p.res_16 + row * p.ld_res + col);
  if (p.res2_16 != nullptr) add_res16<BF16>(x, p.res2_16 + row * p.ld_res + col);
  const long long o = row * p.ld_out + col;
  if (p.out_f32 != nullptr) *reinterpret_cast<float4*>(p.out_f32 + o) = x;
  if (p.out_16 != nullptr) store4_16<BF16>(p.out_16 + o, x);
  if (p.out_16_relu != nullptr)
    store4_16<BF16>(p.out_16_relu + o, make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f)));
}

// ------------------------------------------------------------------------------------------------------------------
// Fused-dot modes (HEAD1X1 / HYPER): shared state between the epilogue warps and the HELPER warp (4th pipeline warp).
// A single epilogue warp pays a full L1/L2 round trip for every global access it makes, so everything that touches
// global memory is moved off the critical warps:
//   helper: stage(i)     bias + second-stage weights of tile i -> opnd[i & 1]      (sfull / sempty, one tile ahead)
//           finalize(i)  sum the per-warpgroup partial dot products of tile i, index math, global stores (pfull / pfree)
//   epilogue warp: TMEM -> +bias -> activation -> dot with the staged weights (all from shared memory) -> partials
// ------------------------------------------------------------------------------------------------------------------
struct DotShared {
  uint64_t sfull[2], sempty[2], pfull, pfree;
  float opnd[2][1024];                     // [1 + c2][block_n] floats per buffer: row 0 = bias, rows 1.. = weights
  float part[kMaxEpiGroups][8][128];       // partial dot products [warpgroup][channel][row] (conflict-free both ways)
};
struct DotSharedNone { int unused; };

// grouped weights: index of the weight block / bias vector this tile uses
L4P_DEVICE int tile_group(const GemmKParams& p, const TileCoord& tc) {
  if (p.a_mode == L4P_A_CONV3D) {
    if (p.conv_grp_b <= 0) return 0;
    const int g = tc.b / p.conv_grp_b, gmax = (p.cB - 1) / p.conv_grp_b;   // the padding block of an odd tile count lies beyond cB
    return g < gmax ? g : gmax;
  }
  return p.grp_a_rows > 0 ? (int)(((long long)tc.m_blk * p.m_stride) / p.grp_a_rows) : 0;
}

struct RowInfo {
  long long row;  // logical output row
  bool ok;
  int cb, ct, ch, cw;
};
L4P_DEVICE RowInfo row_info(const GemmKParams& p, const TileCoord& tc, const int r) {
  RowInfo ri;
  ri.cb = ri.ct = ri.ch = ri.cw = 0;
  if (p.a_mode == L4P_A_CONV3D) {
    const int wl = r % p.bW;
    const int hl = (r / p.bW) % p.bH;
    const int tl = r / (p.bW * p.bH);
    ri.ct = tc.t0 + tl; ri.ch = tc.h0 + hl; ri.cw = tc.w0 + wl; ri.cb = tc.b;
    ri.ok = (ri.ct < p.cT) && (ri.ch < p.cH) && (ri.cw < p.cW) && (ri.cb < p.cB);
    ri.row = (((long long)ri.cb * p.cT + ri.ct) * p.cH + ri.ch) * p.cW + ri.cw;
  } else {
    ri.row = (long long)tc.m_blk * p.m_stride + r;
    ri.ok = ri.row < p.M && r < p.m_stride;
  }
  return ri;
}

// helper warp: bias and second-stage weights of one tile -> opnd buffer (all loads in flight at once)
template <int STORE>
L4P_DEVICE void dot_stage(const GemmKParams& p, const TileCoord& tc, float* opnd, const int lane) {
  const int c2 = p.c2, bn = p.block_n;
  const int n0 = tc.n_blk * bn;
  const float* wsrc;   // row c of the second-stage weights of this tile starts at wsrc + c * wld
  int wld;
  if constexpr (STORE == L4P_STORE_HYPER) {
    const unsigned grp = p.fd_rpg.div((unsigned)(tc.m_blk * kBlockM));
    wsrc = p.w2 + (long long)grp * (c2 * p.ctCout);
    wld = p.ctCout;
  } else {
    wsrc = p.w2 + n0;
    wld = p.N;
  }
  const int nb4 = bn >> 2;  // float4 per row (<= 64: two per lane)
  // rows 0..4 first (bias + up to 4 weight rows: everything the mask decoder needs), then rows 5..8 if present;
  // every pass has all its loads in flight before the first store
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && c2 <= 4) break;
    float4 sv[5][2];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int rw = pass * 5 + j;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c4 = lane + 32 * h;
        sv[j][h] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rw <= c2 && rw < 9 && c4 < nb4 && n0 + 4 * c4 < p.N) {
          if (rw == 0) { if (p.bias != nullptr) sv[j][h] = *reinterpret_cast<const float4*>(p.bias + n0 + 4 * c4); }
          else sv[j][h] = *reinterpret_cast<const float4*>(wsrc + (long long)(rw - 1) * wld + 4 * c4);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int rw = pass * 5 + j;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c4 = lane + 32 * h;
        if (rw <= c2 && rw < 9 && c4 < nb4) reinterpret_cast<float4*>(opnd + rw * bn)[c4] = sv[j][h];
      }
    }
  }
}

// helper warp: reduce the warpgroups' partials of one tile and store the results (4 rows per lane). A lone warp runs
// dependent integer code at ~6 cycles per instruction, so everything tile-uniform is hoisted and the per-row index
// math uses the host-prepared multiply-shift divisions.
template <int STORE, int GROUPS>
L4P_DEVICE void dot_finalize(const GemmKParams& p, const TileCoord& tc, const DotShared& ds, const uint32_t pfree_bar,
                             const int lane) {
  const int c2 = p.c2;
  float acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = lane + 32 * k;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = 0.f;
      if (c < c2) {
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) a += ds.part[g][c][r];
      }
      acc[k][c] = a;
    }
  }
  mbar_arrive(pfree_bar);  // the partials are in registers: the epilogue warps may overwrite them
  if constexpr (STORE == L4P_STORE_HYPER) {
    // rows are input voxels (g,t,h,w) of the [cB,cT,cH,cW] grid (matrix mode); this N tile = tap (kt,kh,kw)
    const int tapi = tc.n_blk;
    const int kw = tapi % p.sW, kh = (tapi / p.sW) % p.sH, kt = tapi / (p.sW * p.sH);
    const int oH = p.cH * p.sH, oW = p.cW * p.sW;
    const long long plane = (long long)(p.cT * p.sT) * oH * oW;
    const int tapoff = (kt * oH + kh) * oW + kw;
    const uint32_t row0 = (uint32_t)tc.m_blk * kBlockM;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t row = row0 + (uint32_t)(lane + 32 * k);
      if (row >= (uint32_t)p.M) continue;
      uint32_t q_, w_, h_, t_;
      p.fd_cW.divmod(row, q_, w_);
      p.fd_cH.divmod(q_, q_, h_);
      p.fd_cT.divmod(q_, q_, t_);
      float* dst = p.out_f32 + ((long long)q_ * c2 * plane + (long long)((int)(t_ * p.sT * oH + h_ * p.sH) * oW + (int)(w_ * p.sW) + tapoff));
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < c2) dst[c * plane] = acc[k][c];
    }
  } else {
    // conv mode: the tile is a (bT,bH,bW) voxel box
    const long long plane = (long long)p.cT * p.cH * p.cW;
    float b2v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) b2v[c] = c < c2 ? p.b2[c] : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t q_, wl, hl, tl;
      p.fd_bW.divmod((uint32_t)(lane + 32 * k), q_, wl);
      p.fd_bH.divmod(q_, tl, hl);
      const int ct = tc.t0 + (int)tl, ch = tc.h0 + (int)hl, cw = tc.w0 + (int)wl;
      if (!(ct < p.cT && ch < p.cH && cw < p.cW && tc.b < p.cB)) continue;
      float* dst = p.out_f32 + ((long long)tc.b * c2 * plane + (long long)((ct * p.cH + ch) * p.cW + cw));
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < c2) {
          float o = acc[k][c] + b2v[c];
          if (p.exp_out) o = expf(o);
          dst[c * plane] = o;
        }
      }
    }
  }
}

// helper warp main loop. `next(tc)` yields this CTA's tiles in epilogue order and returns false when exhausted.
template <int EPI, class NextTile>
L4P_DEVICE void dot_helper_loop(const GemmKParams& p, DotShared& ds, const int lane, NextTile next) {
  constexpr int STORE = epi_store(EPI);
  TileCoord tc, prev;
  bool have = next(tc);
  for (int i = 0;; ++i) {
    if (have) {
      const int b = i & 1;
      HELPER_STAMP();
      mbar_wait(smem_u32(&ds.sempty[b]), (((uint32_t)i >> 1) & 1u) ^ 1u);
      HELPER_STAMP();
      dot_stage<STORE>(p, tc, ds.opnd[b], lane);
      mbar_arrive(smem_u32(&ds.sfull[b]));
      HELPER_STAMP();
    }
    if (i >= 1) {
      mbar_wait(smem_u32(&ds.pfull), (uint32_t)(i - 1) & 1u);
      HELPER_STAMP();
      dot_finalize<STORE, epi_groups(EPI)>(p, prev, ds, smem_u32(&ds.pfree), lane);
      HELPER_STAMP();
    }
    if (!have) break;
    prev = tc;
    have = next(tc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Epilogue. One 128-row x block_n accumulator tile: TMEM -> registers -> bias / activation / residual -> global memory.
// Shared by the 1-CTA and the 2-CTA (cta_group::2) kernels; `release` hands the accumulator stage back to the MMA warp.
//
// tcgen05.ld gives every thread ONE ROW of the tile (lane = row), which is the worst possible shape for global memory:
// a warp-wide 16-byte access touches 32 different rows. The store modes that write the tile out (ROWMAJOR / QKV /
// CONVT) therefore transpose each 32-row x 32-column fp32 sub-tile through a per-warp 4 KiB swizzled staging buffer:
//   phase A (lane = row)          raw accumulators -> 8 x st.shared.v4
//   phase B (8 lanes = one row)   ld.shared.v4 of 4 consecutive columns -> bias, activation, residual -> coalesced
//                                 128 B (fp32) / 64 B (16-bit) row segments, 4 rows per instruction
// Residuals are fetched in the phase-B shape too (coalesced) one chunk ahead, so their latency hides behind the
// previous chunk; the first fetch is issued before the tile's accumulator is even complete.
// The fused-dot modes (HEAD1X1 / HYPER) keep the row-per-thread form: they reduce along the row and store ~nothing.
// ------------------------------------------------------------------------------------------------------------------
template <bool BF16, int EPI, class Release>
L4P_DEVICE void epilogue_tile(const GemmKParams& p, const TileCoord& tc, const int q4, const int lane, const int egrp,
                              const uint32_t tfull_bar, const uint32_t tfull_phase, const uint32_t t_acc, void* dot_shared,
                              const int tile_ord, const uint32_t stage, Release release) {
  constexpr int STORE = epi_store(EPI);
  constexpr bool GEN = (EPI & EPI_GENERIC) != 0;
  constexpr int kEpiGroups = epi_groups(EPI);
  const int r = q4 * 32 + lane;  // row inside the tile
  const int n0 = tc.n_blk * p.block_n;
  const uint32_t t_addr = t_acc + ((uint32_t)(q4 * 32) << 16);

  if constexpr (STORE == L4P_STORE_ROWMAJOR || STORE == L4P_STORE_QKV || STORE == L4P_STORE_CONVT || STORE == kStoreSplitK) {
    // ---------------------------------------------------------------- transposed (coalesced) store modes
    const RowInfo ri = row_info(p, tc, r);
    const long long row = ri.row;
    const bool row_ok = ri.ok;
    const bool res32 = GEN ? (p.res_f32 != nullptr) : ((EPI & EPI_RES32) != 0);
    const bool res16 = GEN ? (p.res_16 != nullptr || p.res2_16 != nullptr) : ((EPI & EPI_RES16) != 0);
    const bool out32 = GEN ? (p.out_f32 != nullptr) : ((EPI & EPI_OUT32) != 0);
    const bool out16 = GEN ? (p.out_16 != nullptr) : ((EPI & EPI_OUT16) != 0);
    const bool out16r = GEN ? (p.out_16_relu != nullptr) : ((EPI & EPI_OUT16R) != 0);
    const bool has_res = STORE == L4P_STORE_ROWMAJOR && (res32 || res16);

    // per-row indices, gathered into the phase-B shape: iteration `it` of a lane works on row it*4 + lane/8
    int my_o, my_r = 0;
    if constexpr (STORE == kStoreSplitK) {
      my_o = (int)row;
    } else if constexpr (STORE == L4P_STORE_ROWMAJOR) {
      my_o = (int)row;
      my_r = p.res_row_mod > 0 ? (int)(row % p.res_row_mod) : (int)row;
    } else if constexpr (STORE == L4P_STORE_QKV) {
      const int bidx = (int)(row / p.tokens);
      my_o = bidx * p.heads * p.tokens + (int)(row - (long long)bidx * p.tokens);  // row of head 0 in the [B,H,N,dpad] view
    } else {
      long long rr = row;
      const int w_ = (int)(rr % p.cW); rr /= p.cW;
      const int h_ = (int)(rr % p.cH); rr /= p.cH;
      const int t_ = (int)(rr % p.cT); rr /= p.cT;
      my_o = (int)(((rr * (p.cT * p.sT) + (long long)t_ * p.sT) * (p.cH * p.sH) + (long long)h_ * p.sH) * (p.cW * p.sW) +
                   (long long)w_ * p.sW);  // output voxel of tap (0,0,0)
    }
    const int sub = lane & 7, rgrp = lane >> 3;
    int orow[8], rrow[8];
    uint32_t okm = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int src = it * 4 + rgrp;
      orow[it] = __shfl_sync(0xffffffffu, my_o, src);
      rrow[it] = res32 ? __shfl_sync(0xffffffffu, my_r, src) : 0;
      okm |= (__shfl_sync(0xffffffffu, row_ok ? 1u : 0u, src) & 1u) << it;
    }
    const int ncols = min(p.block_n, p.N - n0);  // valid columns of this tile
    const int bias_grp = p.conv_grp_b > 0 ? tile_group(p, tc) * p.N : 0;   // grouped conv: one bias vector per group
    const int D = p.heads * p.head_dim;
    const int ld_out = p.ld_out, ld_res = p.ld_res;
    const float* const res_f32 = p.res_f32;
    const uint16_t* const res_16 = p.res_16;
    const uint16_t* const res2_16 = p.res2_16;
    float* const out_f32 = p.out_f32;
    uint16_t* const out_16 = p.out_16;
    uint16_t* const out_16_relu = p.out_16_relu;
    const int act_rt = p.act;

    // residual of the 4-column group (n0 + c0 + 4*sub) of phase-B row `it`. The loads stay RAW in registers (fp32 words /
    // packed 16-bit pairs) until the chunk that consumes them: converting a 16-bit residual at fetch time puts a dependent
    // instruction right behind every load and serialises the whole prefetch (measured: res_16 -> out_16 at M = 262144,
    // N = 1408, K = 704 took 2054 us against 794 us with an fp32 residual of twice the bytes).
    struct ResRaw { float4 f; uint2 h, h2; };
    auto fetch_res = [&](const int c0, const int it, const bool want = true) -> ResRaw {
      const int colg = c0 + sub * 4;
      ResRaw a;
      a.f = make_float4(0.f, 0.f, 0.f, 0.f);
      a.h = make_uint2(0u, 0u);   // +0.0 in both 16-bit formats
      a.h2 = make_uint2(0u, 0u);
      const bool ok = want && colg < ncols && ((okm >> it) & 1u);
      if constexpr (STORE == L4P_STORE_ROWMAJOR) {
        const long long cb = n0 + colg;   // one IMAD.WIDE per address: row * ld + column
        if (res32) a.f = ldg_f32x4_if(res_f32 + ((long long)rrow[it] * ld_res + cb), ok);
        if (res16) {
          const long long o = (long long)orow[it] * ld_res + cb;
          if (res_16 != nullptr) a.h = ldg_res16x4_if(res_16 + o, ok);
          if (res2_16 != nullptr) a.h2 = ldg_res16x4_if(res2_16 + o, ok);
        }
      }
      return a;
    };
    auto add_res = [&](float4& x, const ResRaw& a) {
      if (res32) { x.x += a.f.x; x.y += a.f.y; x.z += a.f.z; x.w += a.f.w; }
      if (res16) {
        const float2 l0 = unpack2<BF16>(a.h.x), h0 = unpack2<BF16>(a.h.y);
        const float2 l1 = unpack2<BF16>(a.h2.x), h1 = unpack2<BF16>(a.h2.y);
        // same association as before: (res_16 + res2_16) is formed first, then added to the accumulator
        x.x += l0.x + l1.x; x.y += l0.y + l1.y; x.z += h0.x + h1.x; x.w += h0.y + h1.y;
      }
    };

    ResRaw rcur[8];
    int c0 = egrp * kEpiChunk;
    if (has_res && c0 < ncols) {  // in flight while the MMA warp finishes the tile
#pragma unroll
      for (int it = 0; it < 8; ++it) rcur[it] = fetch_res(c0, it);
    }

    mbar_wait(tfull_bar, tfull_phase);
    tc_fence_after();
    EPI_STAMP();   // tuning builds: [acc ready | per chunk: tmem loaded, staged, staging read, stored]

    for (; c0 < p.block_n; c0 += kEpiChunk * kEpiGroups) {
      uint32_t raw[32];
      __syncwarp();  // tcgen05.ld is warp-collective; also orders the previous chunk's staging reads before new writes
      tmem_ld32(t_addr + (uint32_t)c0, raw);
      tmem_ld_wait();
      EPI_STAMP();
      if (c0 >= ncols) continue;  // uniform across the CTA
      const int col0 = n0 + c0;

      if constexpr (STORE == L4P_STORE_QKV) {
        if (col0 + kEpiChunk > 2 * D) {
          // V section (or a chunk straddling into it): V^T is written keys-contiguous, which the row-per-thread shape
          // already coalesces (32 lanes = 32 consecutive keys)
          if (row_ok) {
            const long long bidx = row / p.tokens;
            const int tok = (int)(row - bidx * p.tokens);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = col0 + g * 8;
              if (c0 + g * 8 < ncols) {
                const int sct = col / D;
                const int rem = col - sct * D;
                const int h = rem / p.head_dim;
                const int e = rem - h * p.head_dim;
                const long long bh = bidx * p.heads + h;
                float vv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = __uint_as_float(raw[g * 8 + i]) + (p.bias ? p.bias[col + i] : 0.f);
                if (sct < 2) {
                  uint16_t* dst = (sct == 0 ? p.q : p.k) + (bh * p.tokens + tok) * p.head_dim_pad + e;
                  *reinterpret_cast<uint4*>(dst) = make_uint4(pack2<BF16>(vv[0], vv[1]), pack2<BF16>(vv[2], vv[3]),
                                                              pack2<BF16>(vv[4], vv[5]), pack2<BF16>(vv[6], vv[7]));
                } else {
                  uint16_t* dst = p.vt + (bh * p.head_dim_pad + e) * (long long)p.tokens + tok;
#pragma unroll
                  for (int i = 0; i < 8; ++i) dst[(long long)i * p.tokens] = pack1<BF16>(vv[i]);
                }
              }
            }
          }
          continue;
        }
      }

      // phase A: lane = row, 16-byte units XOR-swizzled by the row so that both phases are bank-conflict free
      {
        const uint32_t wbase = stage + (uint32_t)lane * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
        for (int u = 0; u < 8; ++u) sts128(wbase + ((((uint32_t)u) ^ sw) << 4), raw[4 * u], raw[4 * u + 1], raw[4 * u + 2], raw[4 * u + 3]);
      }
      __syncwarp();
      EPI_STAMP();

      // the next chunk's residual is fetched row by row as soon as the current value has been consumed
      const int cn = c0 + kEpiChunk * kEpiGroups;
      const bool more = has_res && cn < ncols;

      // phase B
      const int colg = c0 + sub * 4;
      const bool cok = colg < ncols;
      const int col = n0 + colg;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr && cok) b4 = *reinterpret_cast<const float4*>(p.bias + bias_grp + col);

      // per-lane column decode of the scatter modes (the 4-column group never straddles a head / tap: both are
      // multiples of 8 wide)
      long long coff = 0;     // QKV: h * tokens * dpad + e ; CONVT: tap voxel offset * Cout + co
      uint16_t* sbase = nullptr;
      int srow_ld = 0;
      if constexpr (STORE == L4P_STORE_QKV) {
        const int sct = col / D;
        const int rem = col - sct * D;
        const int h = rem / p.head_dim;
        const int e = rem - h * p.head_dim;
        sbase = (sct == 0 ? p.q : p.k) + ((long long)h * p.tokens * p.head_dim_pad + e);
        srow_ld = p.head_dim_pad;
      } else if constexpr (STORE == L4P_STORE_CONVT) {
        const int tapi = col / p.ctCout;
        const int co = col - tapi * p.ctCout;
        const int kw = tapi % p.sW;
        const int kh = (tapi / p.sW) % p.sH;
        const int kt = tapi / (p.sW * p.sH);
        coff = (((long long)kt * (p.cH * p.sH) + kh) * (p.cW * p.sW) + kw) * p.ctCout + co;
        sbase = p.out_16 + coff;
        srow_ld = p.ctCout;
      }
      const uint32_t rbase = stage + (uint32_t)rgrp * 128u;

      if constexpr (STORE == kStoreSplitK) {
        // K-slice partial sums -> fp32 workspace [M, N]; bias / activation / residuals are applied by the finalize kernel
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const uint32_t rl = (uint32_t)(it * 4 + rgrp);
          const float4 x = lds128(rbase + (uint32_t)it * 512u + ((((uint32_t)sub) ^ (rl & 7u)) << 4));
          if (cok && ((okm >> it) & 1u))  // one 16-byte reduction (red.global.add.v4.f32, sm_90+) instead of four scalar ones
            atomicAdd(reinterpret_cast<float4*>(p.splitk_ws + ((long long)orow[it] * p.N + col)), x);
        }
        continue;
      }
      float4 xs[8];  // all staging reads first: the asm memory clobbers would otherwise serialise them with the stores
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint32_t rl = (uint32_t)(it * 4 + rgrp);
        xs[it] = lds128(rbase + (uint32_t)it * 512u + ((((uint32_t)sub) ^ (rl & 7u)) << 4));
      }
      EPI_STAMP();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        float4 x = xs[it];
        x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
        if constexpr (GEN) {
          apply_act2_rt(x.x, x.y, act_rt); apply_act2_rt(x.z, x.w, act_rt);
        } else {
          apply_act2<epi_act(EPI)>(x.x, x.y); apply_act2<epi_act(EPI)>(x.z, x.w);
        }
        if (has_res) {
          add_res(x, rcur[it]);
          rcur[it] = fetch_res(cn, it, more);   // predicate, not a branch: no reconvergence region per row
        }
        if constexpr (STORE == L4P_STORE_ROWMAJOR) {
          const bool ok = cok && ((okm >> it) & 1u);
          const long long o = (long long)orow[it] * ld_out + col;
          if (out32) stg_f32x4_if(out_f32 + o, x, ok);
          if (out16) stg_b32x2_if(out_16 + o, pack2<BF16>(x.x, x.y), pack2<BF16>(x.z, x.w), ok);
          if (out16r)
            stg_b32x2_if(out_16_relu + o, pack2<BF16>(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)), pack2<BF16>(fmaxf(x.z, 0.f), fmaxf(x.w, 0.f)), ok);
        } else {  // QKV head-major scatter / CONVT pixel shuffle
          stg_b32x2_if(sbase + (long long)orow[it] * srow_ld, pack2<BF16>(x.x, x.y), pack2<BF16>(x.z, x.w), cok && ((okm >> it) & 1u));
        }
      }
      EPI_STAMP();
    }
    // accumulator stage drained -> hand TMEM back to the MMA warp
    tc_fence_before();
    release();
    return;
  } else {
    // ------------------------------------------------------------------ fused-dot modes (row per thread)
    // bias and second-stage weights come from the helper warp's shared-memory staging (see DotShared): no global
    // access on this warp's critical path
    DotShared& ds = *reinterpret_cast<DotShared*>(dot_shared);
    const int c2 = p.c2;
    const int act_rt = p.act;
    const int bn = p.block_n;
    const int b = tile_ord & 1;
    const uint32_t opnd = smem_u32(ds.opnd[b]);
    mbar_wait(smem_u32(&ds.sfull[b]), ((uint32_t)tile_ord >> 1) & 1u);
    EPI_STAMP();

    mbar_wait(tfull_bar, tfull_phase);
    tc_fence_after();
    EPI_STAMP();

    float head_acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) head_acc[c] = 0.f;

    for (int c0 = egrp * 16; c0 < bn; c0 += 16 * kEpiGroups) {
      uint32_t raw[16];
      __syncwarp();  // tcgen05.ld is warp-collective
      tmem_ld16(t_addr + (uint32_t)c0, raw);
      tmem_ld_wait();
      EPI_STAMP();
      if (n0 + c0 >= p.N) continue;  // uniform across the CTA
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 bv = lds128(opnd + (uint32_t)(c0 + i) * 4u);
        v[i] = __uint_as_float(raw[i]) + bv.x; v[i + 1] = __uint_as_float(raw[i + 1]) + bv.y;
        v[i + 2] = __uint_as_float(raw[i + 2]) + bv.z; v[i + 3] = __uint_as_float(raw[i + 3]) + bv.w;
      }
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        if constexpr (GEN) apply_act2_rt(v[i], v[i + 1], act_rt);
        else apply_act2<epi_act(EPI)>(v[i], v[i + 1]);
      }
      // HYPER: dot with the per-query hyper-network vectors; HEAD1X1: the tiny second conv
#pragma unroll
      for (int c = 0; c < (STORE == L4P_STORE_HYPER ? 4 : 8); ++c) {
        if (c < c2) {
          float a[4];  // four independent chains instead of one 16-deep dependent FMA chain
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 wv = lds128(opnd + (uint32_t)((1 + c) * bn + c0 + 4 * i) * 4u);
            a[i] = fmaf(v[4 * i + 3], wv.w, fmaf(v[4 * i + 2], wv.z, fmaf(v[4 * i + 1], wv.y, v[4 * i] * wv.x)));
          }
          head_acc[c] += (a[0] + a[1]) + (a[2] + a[3]);
        }
      }
    }
    // accumulator stage drained -> hand TMEM back to the MMA warp; the operand buffer back to the helper
    EPI_STAMP();
    tc_fence_before();
    release();
    mbar_arrive(smem_u32(&ds.sempty[b]));

    // per-warpgroup partial dot products -> helper warp (which has consumed the previous tile's partials)
    mbar_wait(smem_u32(&ds.pfree), ((uint32_t)tile_ord & 1u) ^ 1u);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < c2) ds.part[egrp][c][r] = head_acc[c];
    mbar_arrive(smem_u32(&ds.pfull));
    EPI_STAMP();
  }
}

template <bool BF16, int EPI>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmKParams p) {
  constexpr int kEpiGroups = epi_groups(EPI);
  constexpr int kWarpProducer = 4 * kEpiGroups, kWarpMma = kWarpProducer + 1, kWarpAlloc = kWarpProducer + 2;
  constexpr bool kRepartition = kEpiGroups == 2;  // setmaxnreg: 96 registers for the pipeline warps, 200 for the epilogue
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  constexpr bool kFusedDot = epi_store(EPI) == L4P_STORE_HEAD1X1 || epi_store(EPI) == L4P_STORE_HYPER;
  typedef typename std::conditional<kFusedDot, DotShared, DotSharedNone>::type DotSh;
  __shared__ __align__(16) DotSh dot_sh;  // fused-dot modes: operand staging + partials shared with the helper warp

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.block_n * 128u;
  const uint32_t kslabs = p.k128 ? 2u : 1u;
  const uint32_t stage_bytes = kslabs * (kABytes + b_bytes);   // [A slab 0][A slab 1][B slab 0][B slab 1]
  const int num_tiles = p.tiles_m * p.tiles_n * p.split_k;  // work units: (tile, K slice); split_k == 1 outside split-K mode

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
    if constexpr (kFusedDot) {
      DotShared& ds = reinterpret_cast<DotShared&>(dot_sh);
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&ds.sfull[s]), 32);
        mbar_init(smem_u32(&ds.sempty[s]), 128 * kEpiGroups);
      }
      mbar_init(smem_u32(&ds.pfull), 128 * kEpiGroups);
      mbar_init(smem_u32(&ds.pfree), 32);
    }
    fence_mbar_init();
  }
  FINE_INIT();
  if (warp == kWarpAlloc) tmem_alloc(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  GRID_STAMP(0);
  pdl_launch_dependents();  // single-wave persistent grid: the next kernel's CTAs may queue up behind ours right away
  pdl_wait();               // everything above overlapped the previous kernel's tail; from here on we touch its outputs
  GRID_STAMP(1);

  if (warp == kWarpProducer) {
    // ------------------------------------------------------------------ TMA producer
    if constexpr (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    int stage = 0, pg = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      if (lane == 0) {
        const TileCoord tc = decode_tile(p, tile / p.split_k);
        const int split = tile % p.split_k;
        const int kb0 = (int)((long long)split * p.num_kb / p.split_k), kb1 = (int)((long long)(split + 1) * p.num_kb / p.split_k);
        // filter-tap walk (cb fastest, then dw, dh, dt) kept as counters: no divisions in the single-thread hot loop
        int cb = 0, dw = -(p.kW / 2), dh = -(p.kH / 2), dt = -(p.kT / 2);
        if (p.a_mode == L4P_A_CONV3D && kb0 > 0) {
          const int tapi = kb0 / p.cblocks;
          cb = kb0 - tapi * p.cblocks;
          dw = tapi % p.kW - p.kW / 2;
          dh = (tapi / p.kW) % p.kH - p.kH / 2;
          dt = tapi / (p.kW * p.kH) - p.kT / 2;
        }
        if (p.k128) {   // matrix mode, two k-blocks per stage (split_k == 1)
          for (int kb = 0; kb < p.num_kb; kb += 2) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_full[stage]);
            mbar_expect_tx(full, stage_bytes);
            tma_load_3d(smem_base + stage * stage_bytes, &tmA, full, 0, tc.m_blk * p.m_stride, kb);
            GEMM_STAMP(0, pg); ++pg;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        } else
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          const uint32_t sa = smem_base + stage * stage_bytes;
          mbar_expect_tx(full, stage_bytes);   // the bytes of BOTH operands (the B warp's load may even land first)
          if (p.a_mode == L4P_A_MATRIX) {
            tma_load_2d(sa, &tmA, full, kb * kBlockK, tc.m_blk * p.m_stride);
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
          GEMM_STAMP(0, pg); ++pg;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
      __syncwarp();
    }
  } else if (warp == kWarpAlloc) {
    // ------------------------------------------------------------------ TMA producer of the B (weight) operand
    // One thread needs ~250 cycles per ring iteration for wait + expect_tx + one TMA issue and ~490 with two loads and the tap walk
    // (tools/ubench/tma_issue_bench.cu, tools/pair_n_sweep.py: every tile width from 64 to 256 took 494 cycles per k-block), which
    // capped all tiles narrower than 256 columns; the two operands are therefore issued by two warps. Both wait on the same
    // empty barrier; the A warp arms the full barrier with the bytes of both loads.
    if constexpr (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile / p.split_k);
        const int split = tile % p.split_k;
        const int kb0 = (int)((long long)split * p.num_kb / p.split_k), kb1 = (int)((long long)(split + 1) * p.num_kb / p.split_k);
        const int n0 = tc.n_blk * p.block_n + tile_group(p, tc) * p.grp_b_rows;
        if (p.k128) {
          for (int kb = 0; kb < p.num_kb; kb += 2) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            tma_load_3d(smem_base + stage * stage_bytes + 2 * kABytes, &tmB, smem_u32(&bar_full[stage]), 0, n0, kb);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        } else
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          tma_load_2d(smem_base + stage * stage_bytes + kABytes, &tmB, smem_u32(&bar_full[stage]), kb * kBlockK, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kWarpMma) {
    // ------------------------------------------------------------------ UMMA issuer
    // whole warp walks the pipeline (warp-uniform control flow), one elected lane issues
    if constexpr (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_f16(BF16, kBlockM, (uint32_t)p.block_n);
    constexpr uint32_t hi128 = umma_desc_hi(128, 2);
    const uint32_t a_lo0 = umma_desc_lo(smem_base), b_lo0 = umma_desc_lo(smem_base + kslabs * kABytes);
    const uint32_t stage_step = stage_bytes >> 4;
    int stage = 0, mg = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
      const int split = tile % p.split_k;
      const int kb0 = (int)((long long)split * p.num_kb / p.split_k), kb1 = (int)((long long)(split + 1) * p.num_kb / p.split_k);
      if (p.k128) {
        // two k-blocks per stage: slab 1 sits one A tile / one B tile behind slab 0; an odd tail stage holds one valid slab
        for (int kb = 0; kb < p.num_kb; kb += 2) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          if (leader) {
            GEMM_STAMP(1, mg);
            const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_step, b_lo = b_lo0 + (uint32_t)stage * stage_step;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_ss(d_tmem, umma_desc_make(a_lo + 2 * k, hi128), umma_desc_make(b_lo + 2 * k, hi128), idesc, (kb | k) != 0 ? 1u : 0u);
            if (kb + 1 < p.num_kb) {
              const uint32_t a_l1 = a_lo + (kABytes >> 4), b_l1 = b_lo + (b_bytes >> 4);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_ss(d_tmem, umma_desc_make(a_l1 + 2 * k, hi128), umma_desc_make(b_l1 + 2 * k, hi128), idesc, 1u);
            }
            umma_commit(smem_u32(&bar_empty[stage]));
            if (kb + 2 >= p.num_kb) umma_commit(smem_u32(&bar_tfull[acc]));
          }
          __syncwarp();
          ++mg;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      } else
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        if (leader) {
          GEMM_STAMP(1, mg);
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_step, b_lo = b_lo0 + (uint32_t)stage * stage_step;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle row: +2 in the (addr >> 4) field
            umma_ss(d_tmem, umma_desc_make(a_lo + 2 * k, hi128), umma_desc_make(b_lo + 2 * k, hi128), idesc,
                    (kb > kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[stage]));
          if (kb == kb1 - 1) umma_commit(smem_u32(&bar_tfull[acc]));
        }
        __syncwarp();
        ++mg;
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp < 4 * kEpiGroups) {
    // ------------------------------------------------------------------ epilogue
    if constexpr (kRepartition) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");  // 32-column sub-tile + prefetched residuals
    const int q4 = warp & 3;  // TMEM lane quarter owned by this warp
    const int egrp = warp >> 2;  // epilogue warpgroup: interleaved 16-column chunks
    int acc = 0, eg = 0;
    uint32_t acc_phase = 0;
    if (threadIdx.x == 0) GEMM_STAMP(2, 511);  // kernel-start reference
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile / p.split_k);
      const uint32_t tempty = smem_u32(&bar_tempty[acc]);
      epilogue_tile<BF16, EPI>(p, tc, q4, lane, egrp, smem_u32(&bar_tfull[acc]), acc_phase, tmem_base + (uint32_t)acc * kAccCols,
                          &dot_sh, eg, smem_base + (uint32_t)p.stages * stage_bytes + (uint32_t)warp * kEpiStageBytes,
                          [&]() { if (threadIdx.x == 0) GEMM_STAMP(2, 2 * eg); mbar_arrive(tempty); });
      if (threadIdx.x == 0) GEMM_STAMP(2, 2 * eg + 1);
      ++eg;
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (kFusedDot && warp == kWarpAlloc + 1) {
    // ------------------------------------------------------------------ fused-dot helper (operand staging + finalisation)
    if constexpr (kFusedDot) {
      int tile = blockIdx.x;
      dot_helper_loop<EPI>(p, reinterpret_cast<DotShared&>(dot_sh), lane, [&](TileCoord& tc) {
        if (tile >= num_tiles) return false;
        tc = decode_tile(p, tile / p.split_k);
        tile += gridDim.x;
        return true;
      });
    }
  } else {
    if constexpr (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
  }

  tc_fence_before();
  __syncthreads();
  FINE_FLUSH();
  GRID_STAMP(2);
  if (warp == kWarpAlloc) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// 2-CTA variant: a CTA pair (cluster of 2, same TPC) owns a 256-row x block_n tile. Each CTA stages its own 128 rows
// of A and HALF of the B tile per k-block (32 KiB instead of 48 KiB per 128 output rows: half the L2 -> SM traffic of the
// B operand), the leader issues tcgen05.mma.cta_group::2 (M=256) which reads both CTAs' shared memory and
// writes each CTA's half of the accumulator into that CTA's own TMEM; both epilogues run independently.
//   full[s]    lives in the leader: 1 arrival (leader's expect_tx of BOTH CTAs' bytes) + the bytes of all four TMA loads
//   empty[s]   one per CTA, released by a multicast tcgen05.commit
//   tfull[a]   one per CTA (multicast commit);   tempty[a] in the leader: all epilogue threads of both CTAs arrive
// ------------------------------------------------------------------------------------------------------------------
template <bool BF16, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm_threads(EPI), 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  constexpr int kEpiGroups = epi_groups(EPI);
  constexpr int kWarpProducer = 4 * kEpiGroups, kWarpMma = kWarpProducer + 1, kWarpAlloc = kWarpProducer + 2;
  constexpr bool kRepartition = kEpiGroups == 2;  // setmaxnreg: 96 registers for the pipeline warps, 200 for the epilogue
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  constexpr bool kFusedDot = epi_store(EPI) == L4P_STORE_HEAD1X1 || epi_store(EPI) == L4P_STORE_HYPER;
  typedef typename std::conditional<kFusedDot, DotShared, DotSharedNone>::type DotSh;
  __shared__ __align__(16) DotSh dot_sh;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t half_n = (uint32_t)p.block_n / 2;
  const uint32_t b_bytes = half_n * 128u;
  const uint32_t kslabs = p.k128 ? 2u : 1u;
  const uint32_t stage_bytes = p.a_halo ? (uint32_t)((p.bH + 2) * p.bW) * 128u + 3u * b_bytes : kslabs * (kABytes + b_bytes);
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
    if constexpr (kFusedDot) {
      DotShared& ds = reinterpret_cast<DotShared&>(dot_sh);
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&ds.sfull[s]), 32);
        mbar_init(smem_u32(&ds.sempty[s]), 128 * kEpiGroups);
      }
      mbar_init(smem_u32(&ds.pfull), 128 * kEpiGroups);
      mbar_init(smem_u32(&ds.pfree), 32);
    }
    fence_mbar_init();
  }
  FINE_INIT();
  if (warp == kWarpAlloc) tmem_alloc2(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast / peer TMA signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  GRID_STAMP(0);
  pdl_launch_dependents();
  pdl_wait();
  GRID_STAMP(1);

  if (warp == kWarpProducer) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if constexpr (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    if (lane == 0) {
      int stage = 0, pg = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n_blk = tile % p.tiles_n;
        const TileCoord tc = decode_block(p, (tile / p.tiles_n) * 2 + (int)rank, n_blk);
        if (p.a_halo) {
          // Line-halo stages: A = the tile's voxel box grown by one line above and below (bH + 2 lines, shifted by dw in W and dt
          // in T), B = the weights of the three taps (dt, -1..1, dw) of one 64-channel block. The three dh taps read the SAME
          // A box at line offsets 0 / 1 / 2, so every activation byte crosses the L2 -> SM port once per dw shift instead of
          // once per tap (9 instead of 27 times: 24 + 3 x 8 KiB per 768 cycles of UMMA instead of 3 x (16 + 8)), and this
          // thread runs one ring iteration per three k-blocks (the B warp issues the three weight tiles).
          for (int dti = 0; dti < p.kT; ++dti)
            for (int dwi = 0; dwi < 3; ++dwi)
              for (int cb = 0; cb < p.cblocks; ++cb) {
                mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
                const uint32_t full_leader = mapa_shared(smem_u32(&bar_full[stage]), 0);
                const uint32_t sa = smem_base + stage * stage_bytes;
                if (is_leader) mbar_expect_tx(smem_u32(&bar_full[stage]), 2 * stage_bytes);
                tma2_load_5d(sa, &tmA, full_leader, cb * kBlockK, tc.w0 + dwi - 1, tc.h0 - 1, tc.t0 + dti - p.kT / 2, tc.b);
                GEMM_STAMP(0, pg); ++pg;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
              }
          continue;
        }
        if (p.k128) {   // matrix mode, two k-blocks per stage
          for (int kb = 0; kb < p.num_kb; kb += 2) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            if (is_leader) mbar_expect_tx(smem_u32(&bar_full[stage]), 2 * stage_bytes);
            tma2_load_3d(smem_base + stage * stage_bytes, &tmA, mapa_shared(smem_u32(&bar_full[stage]), 0), 0, tc.m_blk * kBlockM, kb);
            GEMM_STAMP(0, pg); ++pg;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          continue;
        }
        int cb = 0, dw = -(p.kW / 2), dh = -(p.kH / 2), dt = -(p.kT / 2);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full_leader = mapa_shared(smem_u32(&bar_full[stage]), 0);
          const uint32_t sa = smem_base + stage * stage_bytes;
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
          GEMM_STAMP(0, pg); ++pg;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kWarpAlloc) {
    // ------------------------------------------------------------------ TMA producer of the B operand (both CTAs; see gemm_kernel)
    if constexpr (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_bytes = p.a_halo ? (uint32_t)((p.bH + 2) * p.bW) * 128u : kABytes;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n_blk = tile % p.tiles_n;
        const TileCoord tc = decode_block(p, (tile / p.tiles_n) * 2 + (int)rank, n_blk);
        const int n0 = n_blk * p.block_n + (int)(rank * half_n) + tile_group(p, tc) * p.grp_b_rows;
        if (p.a_halo) {
          for (int dti = 0; dti < p.kT; ++dti)
            for (int dwi = 0; dwi < 3; ++dwi)
              for (int cb = 0; cb < p.cblocks; ++cb) {
                mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
                const uint32_t full_leader = mapa_shared(smem_u32(&bar_full[stage]), 0);
                const uint32_t sb = smem_base + stage * stage_bytes + a_bytes;
#pragma unroll
                for (int dhi = 0; dhi < 3; ++dhi) {
                  const int kb = ((dti * 3 + dhi) * 3 + dwi) * p.cblocks + cb;
                  tma2_load_2d(sb + (uint32_t)dhi * b_bytes, &tmB, full_leader, kb * kBlockK, n0);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
              }
          continue;
        }
        if (p.k128) {
          for (int kb = 0; kb < p.num_kb; kb += 2) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            tma2_load_3d(smem_base + stage * stage_bytes + 2 * kABytes, &tmB, mapa_shared(smem_u32(&bar_full[stage]), 0), 0, n0, kb);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          continue;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          tma2_load_2d(smem_base + stage * stage_bytes + a_bytes, &tmB, mapa_shared(smem_u32(&bar_full[stage]), 0), kb * kBlockK, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kWarpMma) {
    // ------------------------------------------------------------------ UMMA issuer (leader CTA only)
    if constexpr (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    if (is_leader) {
      const bool leader_lane = elect_one();
      const uint32_t idesc = umma_idesc_f16(BF16, 2 * kBlockM, (uint32_t)p.block_n);
      constexpr uint32_t hi128 = umma_desc_hi(128, 2);
      const uint32_t a_lo0 = umma_desc_lo(smem_base), b_lo0 = umma_desc_lo(smem_base + kslabs * kABytes);
      const uint32_t stage_step = stage_bytes >> 4;
      int stage = 0, mg = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
        if (p.a_halo) {
          // one stage = 3 k-blocks: tap dh reads the A box from line dh on (a whole number of 1024-byte swizzle periods: bW is
          // a multiple of 8) and its own weight tile behind the box
          const uint32_t a_units = (uint32_t)((p.bH + 2) * p.bW) * 8u, line_units = (uint32_t)p.bW * 8u, b_units = b_bytes >> 4;
          const int nst = p.kT * 3 * p.cblocks;
          for (int st = 0; st < nst; ++st) {
            mbar_wait(smem_u32(&bar_full[stage]), phase);
            tc_fence_after();
            if (leader_lane) {
              GEMM_STAMP(1, mg);
              const uint32_t s_lo = a_lo0 + (uint32_t)stage * stage_step;
#pragma unroll
              for (int dhi = 0; dhi < 3; ++dhi) {
                const uint32_t a_lo = s_lo + (uint32_t)dhi * line_units, b_lo = s_lo + a_units + (uint32_t)dhi * b_units;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma2_ss(d_tmem, umma_desc_make(a_lo + 2 * k, hi128), umma_desc_make(b_lo + 2 * k, hi128), idesc,
                           (st | dhi | k) != 0 ? 1u : 0u);
              }
              umma2_commit_mc(smem_u32(&bar_empty[stage]), 3);
              if (st == nst - 1) umma2_commit_mc(smem_u32(&bar_tfull[acc]), 3);
            }
            __syncwarp();
            ++mg;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          continue;
        }
        if (p.k128) {
          for (int kb = 0; kb < p.num_kb; kb += 2) {
            mbar_wait(smem_u32(&bar_full[stage]), phase);
            tc_fence_after();
            if (leader_lane) {
              GEMM_STAMP(1, mg);
              const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_step, b_lo = b_lo0 + (uint32_t)stage * stage_step;
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma2_ss(d_tmem, umma_desc_make(a_lo + 2 * k, hi128), umma_desc_make(b_lo + 2 * k, hi128), idesc, (kb | k) != 0 ? 1u : 0u);
              if (kb + 1 < p.num_kb) {
                const uint32_t a_l1 = a_lo + (kABytes >> 4), b_l1 = b_lo + (b_bytes >> 4);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma2_ss(d_tmem, umma_desc_make(a_l1 + 2 * k, hi128), umma_desc_make(b_l1 + 2 * k, hi128), idesc, 1u);
              }
              umma2_commit_mc(smem_u32(&bar_empty[stage]), 3);
              if (kb + 2 >= p.num_kb) umma2_commit_mc(smem_u32(&bar_tfull[acc]), 3);
            }
            __syncwarp();
            ++mg;
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          continue;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          if (leader_lane) {
            GEMM_STAMP(1, mg);
            const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_step, b_lo = b_lo0 + (uint32_t)stage * stage_step;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma2_ss(d_tmem, umma_desc_make(a_lo + 2 * k, hi128), umma_desc_make(b_lo + 2 * k, hi128), idesc,
                       (kb | k) != 0 ? 1u : 0u);
            umma2_commit_mc(smem_u32(&bar_empty[stage]), 3);
            if (kb == p.num_kb - 1) umma2_commit_mc(smem_u32(&bar_tfull[acc]), 3);
          }
          __syncwarp();
          ++mg;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp < 4 * kEpiGroups) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    if constexpr (kRepartition) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");  // 32-column sub-tile + prefetched residuals
    const int q4 = warp & 3;
    const int egrp = warp >> 2;
    int acc = 0, eg = 0;
    uint32_t acc_phase = 0;
    if (threadIdx.x == 0) GEMM_STAMP(2, 511);  // kernel-start reference
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const TileCoord tc = decode_block(p, (tile / p.tiles_n) * 2 + (int)rank, tile % p.tiles_n);
      const uint32_t tempty_leader = mapa_shared(smem_u32(&bar_tempty[acc]), 0);
      epilogue_tile<BF16, EPI>(p, tc, q4, lane, egrp, smem_u32(&bar_tfull[acc]), acc_phase, tmem_base + (uint32_t)acc * kAccCols,
                          &dot_sh, eg, smem_base + (uint32_t)p.stages * stage_bytes + (uint32_t)warp * kEpiStageBytes,
                          [&]() { if (threadIdx.x == 0) GEMM_STAMP(2, 2 * eg); mbar_arrive_cluster(tempty_leader); });
      if (threadIdx.x == 0) GEMM_STAMP(2, 2 * eg + 1);
      ++eg;
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (kFusedDot && warp == kWarpAlloc + 1) {
    // ------------------------------------------------------------------ fused-dot helper (own CTA's 128 rows)
    if constexpr (kFusedDot) {
      int tile = pair;
      dot_helper_loop<EPI>(p, reinterpret_cast<DotShared&>(dot_sh), lane, [&](TileCoord& tc) {
        if (tile >= num_tiles) return false;
        tc = decode_block(p, (tile / p.tiles_n) * 2 + (int)rank, tile % p.tiles_n);
        tile += num_pairs;
        return true;
      });
    }
  } else {
    if constexpr (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
  }

  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still signal / read this CTA
  FINE_FLUSH();
  GRID_STAMP(2);
  if (warp == kWarpAlloc) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// split-K finalisation: out = epilogue(workspace), workspace re-zeroed for the next split-K launch (4 columns per thread).
// A fused form (the CTA that delivers the last K slice of a tile finalises it, per-tile arrival counters) was measured in
// round 2 and removed: one CTA finalising a 128 x 256 tile sits on the critical path for longer than this all-SM kernel
// takes including its (PDL-overlapped) launch: 1.94 vs 1.74 ms for the 59 low-resolution convolutions of a step.
template <bool BF16>
__global__ void __launch_bounds__(256)
splitk_finalize_kernel(const GemmKParams p) {
  pdl_wait();
  const long long n4 = p.N >> 2;
  const long long total = (long long)p.M * n4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / n4;
    const int col = (int)(i - row * n4) * 4;
    float4* w = reinterpret_cast<float4*>(p.splitk_ws + row * p.N + col);
    const float4 x = *w;
    *w = make_float4(0.f, 0.f, 0.f, 0.f);
    splitk_finalize4<BF16>(p, row, col, x);
  }
}

typedef void (*GemmKernelFn)(const CUtensorMap, const CUtensorMap, const GemmKParams);

// One entry per compiled epilogue configuration: [bf16][pair]
struct GemmKernelSet {
  int epi;
  GemmKernelFn fn[2][2];
};
#define L4P_GEMM_KERNEL_SET(E) \
  { (E), { { gemm_kernel<false, (E)>, gemm2_kernel<false, (E)> }, { gemm_kernel<true, (E)>, gemm2_kernel<true, (E)> } } }

// instance groups (gemm_inst_*.cu), searched in order by gemm.cu
const GemmKernelSet* gemm_instances_a(int* n);
const GemmKernelSet* gemm_instances_b(int* n);
const GemmKernelSet* gemm_instances_c(int* n);
const GemmKernelSet* gemm_instances_d(int* n);
void (*gemm_splitk_finalize(bool bf16))(const GemmKParams);

}  // namespace l4p
