// Host side of the tcgen05 GEMM / implicit-GEMM conv kernels (gemm_kernel.cuh): descriptor validation, TMA tensor maps,
// tile / stage selection, 1-CTA vs CTA-pair choice and the lookup of the compile-time epilogue instance.
#include <cstdlib>

#include "gemm_kernel.cuh"

namespace l4p {

// l4p_gemm_plan(): the same host logic as l4p_gemm with the tensor-map encoding and the launch replaced by a report of the
// decisions taken (N tile, split-K, CTA pairing, stages, grid) - testable without a GPU
struct GemmPlanOut { int block_n, split_k, cta_pair, stages, grid, threads; };
static thread_local GemmPlanOut* t_plan = nullptr;

// compiled epilogue instances live in gemm_inst_{a,b,c,d}.cu (split so that they build in parallel)
static GemmKernelFn find_kernel(int epi, bool bf16, bool pair) {
  typedef const GemmKernelSet* (*Getter)(int*);
  static const Getter getters[] = {gemm_instances_a, gemm_instances_b, gemm_instances_c, gemm_instances_d};
  for (Getter g : getters) {
    int n = 0;
    const GemmKernelSet* set = g(&n);
    for (int i = 0; i < n; ++i)
      if (set[i].epi == epi) return set[i].fn[bf16 ? 1 : 0][pair ? 1 : 0];
  }
  return nullptr;
}

// exact compile-time instance for this descriptor, else the run-time-flag instance of its store mode. `wide` asks for the
// three-epilogue-warpgroup instance of the same configuration (short-K problems); *epi_out = the configuration found.
static GemmKernelFn select_kernel(const l4p_gemm_desc* d, bool pair, bool wide = false, int* epi_out = nullptr) {
  int flags = 0, act = d->act;
  if (d->store_mode == L4P_STORE_ROWMAJOR) {
    if (d->res_f32) flags |= EPI_RES32;
    if (d->res_16 || d->res2_16) flags |= EPI_RES16;
    if (d->out_f32) flags |= EPI_OUT32;
    if (d->out_16) flags |= EPI_OUT16;
    if (d->out_16_relu) flags |= EPI_OUT16R;
  }
  if (d->store_mode == L4P_STORE_QKV) act = L4P_ACT_NONE;
  int epi = epi_make(d->store_mode, act, flags);
  GemmKernelFn fn = nullptr;
  if (wide) {
    fn = find_kernel(epi | EPI_WIDE3, d->bf16 != 0, pair);
    if (fn != nullptr) epi |= EPI_WIDE3;
  }
  if (fn == nullptr) fn = find_kernel(epi, d->bf16 != 0, pair);
  if (fn == nullptr) { epi = epi_make(d->store_mode, 0, EPI_GENERIC); fn = find_kernel(epi, d->bf16 != 0, pair); }
  if (epi_out != nullptr) *epi_out = epi;
  return fn;
}

// Matrix mode with K % 64 == 0: ring stages of TWO 64-wide k-blocks per operand. The operand is viewed as [K / 64][rows][64]
// (the k-block index is the slowest TMA dimension, 128 bytes apart), one 3-D box {64, rows, 2} per operand and stage lands
// slab-major in shared memory: [rows x 128 B of k-block kb][rows x 128 B of k-block kb + 1]. A box that reaches past the last
// k-block is zero-filled (odd k-block counts: the issuer skips the second slab). L4P_GEMM_K128=0 disables (A / B runs).
static bool k128_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("L4P_GEMM_K128");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}
static int make_tmap_k128(CUtensorMap* m, const void* base, uint64_t K, uint64_t rows, uint64_t ld_elems, uint32_t box_rows) {
  const uint64_t dims[3] = {(uint64_t)kBlockK, rows, K / kBlockK};
  const uint64_t strides[2] = {ld_elems * 2, (uint64_t)kBlockK * 2};
  const uint32_t box[3] = {kBlockK, box_rows, 2};
  return host_make_tmap_16b(m, base, 3, dims, strides, box, 128);
}

// N-tile width from a measured cost model (tools/pair_n_sweep.py, tools/blockn_sweep2.py). With the operands issued by two warps and
// 128-wide K stages the 2-CTA kernel is tensor-bound down to 96-column tiles: one k-block of a 256 x bn pair tile costs ~2 bn cycles,
// a tile a fixed ~2300 cycles more (pipeline drain, epilogue hand-over), and the kernel needs ceil(tiles / CTA pairs) rounds of tiles.
// The 1-CTA kernel (128-row tiles) still measures ~480 + 0.4 bn cycles per k-block whatever the width, so it wants the fewest,
// widest tiles. A ragged last N tile (TMA zero fill + column masking) is fine. Examples at M = 2048: N = 1408 -> 9 x 160 (72 pair
// tiles on 74 pairs, one round; 8 x 176 costs 10 % more per k-block), N = 6144 -> 26 x 240 (3 rounds like 24 x 256, 6 % cheaper),
// N = 4224 -> 17 x 256 (2 rounds; 22 x 192 needs 3).
static int pick_block_n(long long N, long long tiles_m, long long num_kb) {
  if (N <= 128) return (int)((N + 15) / 16 * 16);
  const int sms = host_num_sms();
  const int cands[] = {256, 240, 224, 208, 192, 176, 160, 144, 128};
  int dflt = 256;  // the widest tile that divides N ...
  for (int c : cands)
    if (N % c == 0) { dflt = c; break; }
  auto is_pair = [&](int c) {
    const long long pair_tiles = ((tiles_m + 1) / 2) * ((N + c - 1) / c);
    return pair_tiles * 5 >= (sms / 2) * 4 && tiles_m >= 2;   // same rule as the kernel choice below
  };
  auto cost = [&](int c) {
    const long long tn = (N + c - 1) / c;
    const long long pair_tiles = ((tiles_m + 1) / 2) * tn;
    const bool pair = is_pair(c);
    if (pair) return (double)((pair_tiles + sms / 2 - 1) / (sms / 2)) * (2.0 * c * (double)num_kb + 2300.0);
    return (double)((tiles_m * tn + sms - 1) / sms) * ((480.0 + 0.4 * c) * (double)num_kb + 2300.0);
  };
  // ... unless another (ragged) tiling is predicted >= 3 % cheaper (narrower tiles only for the 2-CTA kernel: small problems on
  // the 1-CTA kernel keep their tile count, which the split-K and narrow-tile rules below are tuned on)
  int best = dflt;
  double best_cost = cost(dflt) * 0.97;
  for (int c : cands) {
    if (c == dflt || (c < dflt && !is_pair(c))) continue;
    const double k = cost(c);
    if (k < best_cost) { best_cost = k; best = c; }
  }
  return best;
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

  L4P_REQUIRE(d->a_mode == L4P_A_MATRIX || d->a_mode == L4P_A_CONV3D, L4P_ERR_ARG, "l4p_gemm: a_mode=%d", d->a_mode);
  if (d->a_mode == L4P_A_CONV3D) {
    // validate the box before it is used as a divisor: a zeroed / malformed conv descriptor is an argument error, not a SIGFPE
    L4P_REQUIRE(d->bT > 0 && d->bH > 0 && d->bW > 0 && d->bT * d->bH * d->bW == kBlockM, L4P_ERR_SHAPE,
                "l4p_gemm: conv box %dx%dx%d must have %d voxels", d->bT, d->bH, d->bW, kBlockM);
    L4P_REQUIRE(d->cB > 0 && d->cT > 0 && d->cH > 0 && d->cW > 0 && d->cCin > 0 && d->kT > 0 && d->kH > 0 && d->kW > 0,
                L4P_ERR_SHAPE, "l4p_gemm: conv geometry B=%d T=%d H=%d W=%d Cin=%d k=%dx%dx%d", d->cB, d->cT, d->cH, d->cW,
                d->cCin, d->kT, d->kH, d->kW);
  }

  // grouped convolution (several heads' identical layers in one launch): batch entries per group
  const bool conv_grouped = d->conv_grp_b > 0;
  if (conv_grouped) {
    L4P_REQUIRE(d->a_mode == L4P_A_CONV3D && d->store_mode == L4P_STORE_ROWMAJOR && d->grp_a_rows == 0, L4P_ERR_ARG,
                "l4p_gemm(grouped conv): conv mode with row-major store only");
    L4P_REQUIRE(d->cB % d->conv_grp_b == 0 && (int64_t)(d->cB / d->conv_grp_b) * d->N < (1ll << 31), L4P_ERR_SHAPE,
                "l4p_gemm(grouped conv): cB=%d must be a multiple of conv_grp_b=%d", d->cB, d->conv_grp_b);
  }
  // grouped weights / tile row stride (see l4p_b200.h): matrix mode, row-major store, 1-CTA kernel, no split-K
  const bool grouped = d->grp_a_rows > 0;
  const int m_stride = d->m_stride > 0 ? d->m_stride : kBlockM;
  if (grouped || m_stride != kBlockM) {
    L4P_REQUIRE(d->a_mode == L4P_A_MATRIX && d->store_mode == L4P_STORE_ROWMAJOR && d->cta_pair != 1, L4P_ERR_ARG,
                "l4p_gemm(grouped): matrix mode, row-major store and the 1-CTA kernel only");
    L4P_REQUIRE(m_stride >= 8 && m_stride <= kBlockM && m_stride % 8 == 0, L4P_ERR_SHAPE, "l4p_gemm: m_stride=%d", m_stride);
    if (grouped)
      L4P_REQUIRE(d->grp_a_rows % m_stride == 0 && d->grp_b_rows >= d->N && d->grp_a_rows < (1ll << 31) && d->grp_b_rows < (1ll << 31) &&
                      ((d->M + d->grp_a_rows - 1) / d->grp_a_rows) * d->grp_b_rows < (1ll << 31),
                  L4P_ERR_SHAPE, "l4p_gemm(grouped): grp_a_rows=%lld must be a multiple of m_stride=%d, grp_b_rows=%lld >= N",
                  (long long)d->grp_a_rows, m_stride, (long long)d->grp_b_rows);
  }

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)d->M;
  p.N = (int)d->N;
  p.m_stride = m_stride;
  p.grp_a_rows = (int)d->grp_a_rows;
  p.grp_b_rows = conv_grouped ? (int)d->N : (int)d->grp_b_rows;
  p.conv_grp_b = d->conv_grp_b;
  const long long tm_all = d->a_mode == L4P_A_CONV3D
                               ? (long long)d->cB * ((d->cT + d->bT - 1) / d->bT) * ((d->cH + d->bH - 1) / d->bH) * ((d->cW + d->bW - 1) / d->bW)
                               : (d->M + m_stride - 1) / m_stride;
  p.block_n = d->block_n > 0 ? d->block_n : pick_block_n(d->N, tm_all, (d->K + kBlockK - 1) / kBlockK);
  if (d->block_n <= 0 && d->store_mode == L4P_STORE_HYPER) p.block_n = d->ctCout;  // one tap per tile
  // Few rows and a short K loop (the track head's token-side GEMMs: M = 128 queries x 6 tokens = 768 rows, K <= 2048; the DPT
  // heads' token projections: M = 2048, N <= 1024): one wave
  // of NARROW tiles - the smallest N tile that still gives every tile its own SM - beats wide tiles cut along K (split-K +
  // finalize kernel): 10.4 vs 17.4 us for 768 x 1408 x 1408, 11.3 vs 21.4 us for 768 x 2048 x 1408 inside a dependent chain
  // (tools/small_gemm_sweep.py). Long-K problems (the low-resolution convolutions, K = 6912+) keep split-K: narrow tiles would
  // re-read A once per N tile.
  bool narrow = false;
  if (d->block_n <= 0 && d->a_mode == L4P_A_MATRIX && d->store_mode == L4P_STORE_ROWMAJOR && !grouped && m_stride == kBlockM &&
      tm_all <= 16 && (d->K + kBlockK - 1) / kBlockK <= 32 && d->N >= 64) {
    const long long sms = host_num_sms();
    int bn = 32;
    while (bn < 256 && tm_all * ((d->N + bn - 1) / bn) > sms) bn += 16;
    if (bn <= 128) {   // wider one-wave tiles (M = 2048 with N >= 1408: the encoder GEMMs) are better served by the 2-CTA kernel
      p.block_n = bn;
      narrow = true;
    }
  }
  // Mid-size 3x3 convolutions (the grouped 8x16x16 / 16x16x16 RefineNet levels: 32..128 row tiles, N = 256): two 128-wide N
  // tiles through the 2-CTA kernel with line-halo stages beat one 256-wide tile cut along K with a finalize kernel
  // (tools/small_conv_sweep.py: 8x16x16 256 -> 256 x3: 22.1 vs 32.6 us; 16x16x16: 37.1 vs 42.8 us; 1024 -> 256: 69.9 vs 76.3 us).
  bool halo_pair_pref = false;
  if (d->block_n <= 0 && d->cta_pair == 0 && d->a_mode == L4P_A_CONV3D && d->store_mode == L4P_STORE_ROWMAJOR && d->kH == 3 &&
      d->kW == 3 && d->bT == 1 && d->bH >= 2 && d->bW % 8 == 0 && tm_all >= 32 && tm_all <= 128 && tm_all % 2 == 0 &&
      d->N >= 256 && d->N % 128 == 0 &&
      (!conv_grouped || ((long long)d->conv_grp_b * ((d->cT + d->bT - 1) / d->bT) * ((d->cH + d->bH - 1) / d->bH) * ((d->cW + d->bW - 1) / d->bW)) % 2 == 0)) {
    p.block_n = 128;
    halo_pair_pref = true;
  }
  // split-K decision (needs the tile and k-block counts up front)
  int split_k = 1;
  {
    const long long tm = tm_all;
    const long long tiles = tm * ((d->N + p.block_n - 1) / p.block_n);
    const long long nkb = d->a_mode == L4P_A_CONV3D ? (long long)d->kT * d->kH * d->kW * (d->cCin / kBlockK) : (d->K + kBlockK - 1) / kBlockK;
    const bool can = d->store_mode == L4P_STORE_ROWMAJOR && d->splitk_ws != nullptr && d->splitk_ws_bytes >= d->M * d->N * 4 &&
                     d->split_k != 1 && d->block_n <= 0 && !grouped && m_stride == kBlockM && !narrow && !halo_pair_pref;
    if (can) {
      if (d->split_k > 1) {
        split_k = d->split_k;
      } else if (tiles <= 48 && nkb >= 16) {
        long long s_ = host_num_sms() / tiles;          // fill the machine ...
        if (s_ > nkb / 4) s_ = nkb / 4;                  // ... with at least 4 k-blocks per slice
        if (s_ >= 2) split_k = (int)s_;
      }
      if (split_k > nkb) split_k = (int)nkb;
      if (split_k < 1) split_k = 1;
    }
  }
  if (d->block_n <= 0 && d->store_mode != L4P_STORE_HEAD1X1 && split_k == 1 && !narrow && !halo_pair_pref) {
    // few output tiles (low-resolution pyramid levels, token-side GEMMs): trade tile width for CTAs so that more
    // than a handful of SMs work on the (long) K loop
    const long long tm = tm_all;
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
    p.tiles_m = (int)tm_all;
    const uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->M};
    const uint64_t strides[1] = {(uint64_t)d->lda * 2};
    const uint32_t box[2] = {kBlockK, kBlockM};
    rc = t_plan ? L4P_OK : host_make_tmap_16b(&tmA, d->a, 2, dims, strides, box, 128);
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
    rc = t_plan ? L4P_OK : host_make_tmap_16b(&tmA, d->a, 5, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  } else {
    return host_set_error(L4P_ERR_ARG, "l4p_gemm: a_mode=%d", d->a_mode);
  }
  {
    // grouped: W holds one [grp_b_rows, ldw] block per group of A rows
    const uint64_t w_rows = grouped ? (uint64_t)((d->M + d->grp_a_rows - 1) / d->grp_a_rows) * (uint64_t)d->grp_b_rows
                            : conv_grouped ? (uint64_t)(d->cB / d->conv_grp_b) * (uint64_t)d->N : (uint64_t)d->N;
    const uint64_t dims[2] = {(uint64_t)d->K, w_rows};
    const uint64_t strides[1] = {(uint64_t)d->ldw * 2};
    const uint32_t box[2] = {kBlockK, (uint32_t)p.block_n};
    rc = t_plan ? L4P_OK : host_make_tmap_16b(&tmB, d->w, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
  }

  // epilogue validation
  p.bias = d->bias;
  p.act = d->act;
  p.store_mode = d->store_mode;
  p.res_f32 = d->res_f32;
  p.res_16 = (const uint16_t*)d->res_16;
  p.res2_16 = (const uint16_t*)d->res2_16;
  p.ld_res = (int)d->ld_res;
  p.res_row_mod = d->res_row_mod;
  p.prof = (long long*)d->prof;
  p.out_f32 = d->out_f32;
  p.out_16 = (uint16_t*)d->out_16;
  p.out_16_relu = (uint16_t*)d->out_16_relu;
  p.ld_out = (int)d->ld_out;
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
      L4P_REQUIRE((1 + d->c2) * p.block_n * 4 <= kEpiStageBytes, L4P_ERR_SHAPE,
                  "l4p_gemm(head1x1): (1+c2)*N = %d floats exceed the %d-byte epilogue staging", (1 + d->c2) * p.block_n, kEpiStageBytes);
      p.w2 = d->w2; p.b2 = d->b2; p.c2 = d->c2; p.exp_out = d->exp_out;
      L4P_REQUIRE((long long)d->cT * d->cH * d->cW < (1ll << 31), L4P_ERR_SHAPE, "l4p_gemm(head1x1): volume too large");
      p.fd_bW = make_fastdiv(d->bW); p.fd_bH = make_fastdiv(d->bH);
      break;
    case L4P_STORE_HYPER:
      L4P_REQUIRE(d->out_f32 && d->w2 && d->c2 >= 1 && d->c2 <= 4, L4P_ERR_ARG, "l4p_gemm(hyper): args");
      L4P_REQUIRE(d->ctCout % 16 == 0 && d->N == (int64_t)d->sT * d->sH * d->sW * d->ctCout, L4P_ERR_SHAPE,
                  "l4p_gemm(hyper): N != sT*sH*sW*Cout");
      L4P_REQUIRE(p.block_n == d->ctCout, L4P_ERR_SHAPE, "l4p_gemm(hyper): block_n must equal Cout (one tap per tile)");
      L4P_REQUIRE(d->rows_per_group % kBlockM == 0 && (1 + d->c2) * p.block_n * 4 <= kEpiStageBytes, L4P_ERR_SHAPE,
                  "l4p_gemm(hyper): rows_per_group=%lld must be a multiple of 128 and (1+c2)*Cout <= 1024", (long long)d->rows_per_group);
      L4P_REQUIRE(d->M == (int64_t)d->cB * d->cT * d->cH * d->cW && d->rows_per_group > 0, L4P_ERR_SHAPE,
                  "l4p_gemm(hyper): M mismatch");
      p.cB = d->cB; p.cT = d->cT; p.cH = d->cH; p.cW = d->cW;
      p.sT = d->sT; p.sH = d->sH; p.sW = d->sW; p.ctCout = d->ctCout;
      p.w2 = d->w2; p.c2 = d->c2; p.rows_per_group = d->rows_per_group;
      L4P_REQUIRE(d->M < (1ll << 31) && d->rows_per_group < (1ll << 31), L4P_ERR_SHAPE, "l4p_gemm(hyper): M too large");
      p.fd_cW = make_fastdiv(d->cW); p.fd_cH = make_fastdiv(d->cH); p.fd_cT = make_fastdiv(d->cT);
      p.fd_rpg = make_fastdiv((int)d->rows_per_group);
      break;
    default:
      return host_set_error(L4P_ERR_ARG, "l4p_gemm: store_mode=%d", d->store_mode);
  }

  const uint32_t stage_bytes = kABytes + (uint32_t)p.block_n * 128u;
  const bool fused_dot = d->store_mode == L4P_STORE_HEAD1X1 || d->store_mode == L4P_STORE_HYPER;
  const int threads = gemm_threads(d->store_mode);  // epi_groups() only looks at the store-mode bits
  // store modes: per-warp transposition buffers behind the ring (dynamic); fused-dot modes: the DotShared block (static)
  const uint32_t epi_bytes = fused_dot ? 0u : (uint32_t)epi_smem_bytes(epi_groups(d->store_mode));
  // 227 KiB per CTA = ring + 1 KiB alignment + epilogue staging + static (barriers, DotShared ~25 KiB in the fused-dot modes)
  const uint32_t kRingBudget = (fused_dot ? 196u : 190u) * 1024u;
  int stages = (int)(kRingBudget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > p.num_kb) stages = p.num_kb < 2 ? 2 : p.num_kb;
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + epi_bytes;

  p.split_k = split_k;
  p.splitk_ws = (float*)d->splitk_ws;
  const int num_tiles = p.tiles_m * p.tiles_n * split_k;
  int grid = host_num_sms();
  if (grid > num_tiles) grid = num_tiles;

  if (split_k > 1) {
    // K slices accumulate into the fp32 workspace; the finalize kernel applies the epilogue and re-zeroes it
    GemmKernelFn ks = find_kernel(epi_make(kStoreSplitK, L4P_ACT_NONE, 0), d->bf16 != 0, false);
    L4P_REQUIRE(ks != nullptr, L4P_ERR_ARG, "l4p_gemm: split-K kernel instance missing");
    if (t_plan) { *t_plan = GemmPlanOut{p.block_n, split_k, 0, p.stages, grid, threads}; return L4P_OK; }
    L4P_CHECK_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRingBudget + 1024 + epi_bytes)));
    L4P_CHECK_CUDA(launch_pdl(ks, dim3(grid), dim3(threads), smem, stream, tmA, tmB, p));
    const long long n4 = (long long)p.M * (p.N / 4);
    long long fgrid = (n4 + 255) / 256;
    if (fgrid > 8ll * host_num_sms()) fgrid = 8ll * host_num_sms();
    L4P_CHECK_CUDA(launch_pdl(gemm_splitk_finalize(d->bf16 != 0), dim3((unsigned)fgrid), dim3(256), 0, stream, p));
    return L4P_OK;
  }

  // CTA pairs (256-row tiles) when the problem is large enough to fill the machine with pair tiles
  const int pair_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;
  const int pairs = host_num_sms() / 2;
  int use_pair = d->cta_pair;  // 0 = auto, 1 = force, -1 = never
  // the pair kernel halves the B traffic per SM: worth it even when it leaves a few pairs idle (M=2048, N=1408: 64 pair
  // tiles on 74 pairs beat 128 single tiles on 148 SMs by 6 % at K=6144, equal at K=1408)
  if (grouped || m_stride != kBlockM) use_pair = -1;
  if (halo_pair_pref) use_pair = 1;
  // grouped conv: the two 128-row blocks of a pair tile share ONE weight tile, so a pair must not straddle two groups
  if (conv_grouped && ((long long)d->conv_grp_b * p.ntT * p.ntH * p.ntW) % 2 != 0) {
    L4P_REQUIRE(use_pair != 1, L4P_ERR_SHAPE, "l4p_gemm(grouped conv): odd tile count per group, the 2-CTA kernel cannot be forced");
    use_pair = -1;
  }
  if (use_pair == 0) use_pair = (pair_tiles * 5 >= pairs * 4 && p.block_n >= 64 && p.tiles_m >= 2) ? 1 : -1;
  if (use_pair == 1) {
    L4P_REQUIRE(p.block_n % 32 == 0 || p.block_n % 16 == 0, L4P_ERR_SHAPE, "l4p_gemm(pair): block_n=%d", p.block_n);
    // B tensor map with the half-tile box
    const uint64_t dims[2] = {(uint64_t)d->K, conv_grouped ? (uint64_t)(d->cB / d->conv_grp_b) * (uint64_t)d->N : (uint64_t)d->N};
    const uint64_t strides[1] = {(uint64_t)d->ldw * 2};
    const uint32_t box[2] = {kBlockK, (uint32_t)(p.block_n / 2)};
    rc = t_plan ? L4P_OK : host_make_tmap_16b(&tmB, d->w, 2, dims, strides, box, 128);
    if (rc != L4P_OK) return rc;
    // Line-halo stages (gemm_kernel.cuh, gemm2_kernel producer): the three in-plane row taps share one A box with bH + 2
    // lines: one ring iteration (1 + 3 TMA loads, one wait / expect_tx / commit) per THREE k-blocks and a third of the activation
    // traffic. Measured 927 -> 1519 TFLOP/s on the 224^2 head convolution; the plain stages were capped by the producer thread's
    // ~490 cycles per k-block (then read as a 64 B/clk L2 -> SM port limit).
    // L4P_CONV_HALO=0 restores one box per tap (tuning / A-B).
    static int halo_env = -1;
    if (halo_env < 0) {
      const char* e = getenv("L4P_CONV_HALO");
      halo_env = (e && atoi(e) == 0) ? 0 : 1;
    }
    const uint32_t halo_a = (uint32_t)((d->bH + 2) * d->bW) * 128u;
    const bool halo = halo_env && d->a_mode == L4P_A_CONV3D && d->kH == 3 && d->kW == 3 && d->bT == 1 && d->bH >= 2 && d->bW % 8 == 0 &&
                      2u * (halo_a + 3u * (uint32_t)(p.block_n / 2) * 128u) <= kRingBudget;
    if (halo) {
      p.a_halo = 1;
      const uint64_t C = (uint64_t)d->cCin;
      const uint64_t dimsA[5] = {C, (uint64_t)d->cW, (uint64_t)d->cH, (uint64_t)d->cT, (uint64_t)d->cB};
      const uint64_t stridesA[4] = {C * 2, C * 2 * d->cW, C * 2 * d->cW * d->cH, C * 2 * d->cW * d->cH * d->cT};
      const uint32_t boxA[5] = {kBlockK, (uint32_t)d->bW, (uint32_t)(d->bH + 2), 1, 1};
      rc = t_plan ? L4P_OK : host_make_tmap_16b(&tmA, d->a, 5, dimsA, stridesA, boxA, 128);
      if (rc != L4P_OK) return rc;
    }
    uint32_t sb2 = halo ? halo_a + 3u * (uint32_t)(p.block_n / 2) * 128u : kABytes + (uint32_t)(p.block_n / 2) * 128u;
    int nst2 = halo ? d->kT * 3 * p.cblocks : p.num_kb;   // ring iterations per tile
    // two k-blocks per stage when three such stages fit (tiles up to 256 x 208): the producer / issuer threads then keep up with
    // tiles narrower than 256 columns (fc2 / proj at N = 176: 407 -> 352 cycles per k-block)
    if (k128_enabled() && d->a_mode == L4P_A_MATRIX && d->K % kBlockK == 0 && p.num_kb >= 4 && 3u * 2u * sb2 <= kRingBudget) {
      p.k128 = 1;
      sb2 *= 2;
      nst2 = (p.num_kb + 1) / 2;
      if (!t_plan) {
        rc = make_tmap_k128(&tmA, d->a, (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda, kBlockM);
        if (rc != L4P_OK) return rc;
        rc = make_tmap_k128(&tmB, d->w, (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldw, (uint32_t)(p.block_n / 2));
        if (rc != L4P_OK) return rc;
      }
    }
    int st2 = (int)(kRingBudget / sb2);
    if (st2 > kMaxStages) st2 = kMaxStages;
    if (st2 > nst2) st2 = nst2 < 2 ? 2 : nst2;
    if (st2 > kMaxStages) st2 = kMaxStages;
    p.stages = st2;
    const size_t smem2 = (size_t)st2 * sb2 + 1024 + epi_bytes;
    int g2 = pairs < pair_tiles ? pairs : pair_tiles;
    GemmKernelFn k2 = select_kernel(d, true);
    L4P_REQUIRE(k2 != nullptr, L4P_ERR_ARG, "l4p_gemm: no kernel instance for store_mode=%d", d->store_mode);
    if (t_plan) { *t_plan = GemmPlanOut{p.block_n, 1, 1, p.stages, 2 * g2, threads}; return L4P_OK; }
    L4P_CHECK_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRingBudget + 1024 + epi_bytes)));
    L4P_CHECK_CUDA(launch_pdl(k2, dim3(2 * g2), dim3(threads), smem2, stream, tmA, tmB, p));
    return L4P_OK;
  }

  // short-K problems (one or two k-blocks per tile: the per-query K = 48 output GEMM of the track head) spend their time in the
  // epilogue: take the instance with three epilogue warpgroups when this configuration has one
  uint32_t stage_bytes1 = stage_bytes;
  if (k128_enabled() && d->a_mode == L4P_A_MATRIX && d->K % kBlockK == 0 && p.num_kb >= 4 && 3u * 2u * stage_bytes <= kRingBudget) {
    p.k128 = 1;
    stage_bytes1 = 2 * stage_bytes;
    int st = (int)(kRingBudget / stage_bytes1);
    const int nst = (p.num_kb + 1) / 2;
    if (st > kMaxStages) st = kMaxStages;
    if (st > nst) st = nst;
    p.stages = st;
    if (!t_plan) {
      const uint64_t w_rows = grouped ? (uint64_t)((d->M + d->grp_a_rows - 1) / d->grp_a_rows) * (uint64_t)d->grp_b_rows : (uint64_t)d->N;
      rc = make_tmap_k128(&tmA, d->a, (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda, kBlockM);
      if (rc != L4P_OK) return rc;
      rc = make_tmap_k128(&tmB, d->w, (uint64_t)d->K, w_rows, (uint64_t)d->ldw, (uint32_t)p.block_n);
      if (rc != L4P_OK) return rc;
    }
  }
  int epi_sel = 0;
  const bool wide = d->store_mode == L4P_STORE_ROWMAJOR && p.num_kb <= 2 && num_tiles >= 4 * grid;
  GemmKernelFn kfn = select_kernel(d, false, wide, &epi_sel);
  L4P_REQUIRE(kfn != nullptr, L4P_ERR_ARG, "l4p_gemm: no kernel instance for store_mode=%d", d->store_mode);
  const int threads1 = gemm_threads(epi_sel);
  const uint32_t epi_bytes1 = fused_dot ? 0u : (uint32_t)epi_smem_bytes(epi_groups(epi_sel));
  const size_t smem1 = (size_t)p.stages * stage_bytes1 + 1024 + epi_bytes1;
  if (t_plan) { *t_plan = GemmPlanOut{p.block_n, 1, 0, p.stages, grid, threads1}; return L4P_OK; }
  const size_t attr1 = (epi_sel & EPI_WIDE3) ? smem1 : (size_t)(kRingBudget + 1024 + epi_bytes1);
  L4P_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attr1));
  L4P_CHECK_CUDA(launch_pdl(kfn, dim3(grid), dim3(threads1), smem1, stream, tmA, tmB, p));
  return L4P_OK;
}

extern "C" int l4p_gemm_plan(const l4p_gemm_desc* d, int* out6) {
  L4P_REQUIRE(d != nullptr && out6 != nullptr, L4P_ERR_ARG, "l4p_gemm_plan: null argument");
  GemmPlanOut po = {0, 0, 0, 0, 0, 0};
  t_plan = &po;
  const int rc = l4p_gemm(d, nullptr);
  t_plan = nullptr;
  if (rc != L4P_OK) return rc;
  out6[0] = po.block_n; out6[1] = po.split_k; out6[2] = po.cta_pair; out6[3] = po.stages; out6[4] = po.grid; out6[5] = po.threads;
  return L4P_OK;
}
