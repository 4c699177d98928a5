// HBM-bound helper kernels around the tensor-core path: tubelet gather for the patch embedding (K1),
// fp32 -> 16-bit casts, channels-last trilinear resampling (K9), strided-conv gather (K7, the one stride-2
// convolution of the DPT reassemble stage). All are coalesced / 16-byte vectorised over the channel axis.
#include "common.cuh"

namespace l4p {

// ------------------------------------------------------------------------------------------------
// K1 gather: rgb [B,C,T,H,W] fp32 -> A [B*nt*nh*nw, C*pt*ph*pw] 16-bit, K ordered (c,dt,dh,dw) which is the
// flattened Conv3d weight order (modeling_finetune.py:268-273), token order t'*nh*nw + h'*nw + w' (:282).
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void patchify_kernel(const float* __restrict__ rgb, uint16_t* __restrict__ out, int B, int C, int T, int H,
                                int W, int pt, int ph, int pw, long long total) {
  const int nt = T / pt, nh = H / ph, nw = W / pw;
  const int K = C * pt * ph * pw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % K);
    long long row = idx / K;
    const int dw = kk % pw;
    const int dh = (kk / pw) % ph;
    const int dt = (kk / (pw * ph)) % pt;
    const int c = kk / (pw * ph * pt);
    const int wq = (int)(row % nw); row /= nw;
    const int hq = (int)(row % nh); row /= nh;
    const int tq = (int)(row % nt); row /= nt;
    const long long b = row;
    const long long src = (((b * C + c) * T + (tq * pt + dt)) * H + (hq * ph + dh)) * (long long)W + (wq * pw + dw);
    out[idx] = pack1<BF16>(rgb[src]);
  }
}

// y = round16(x (+ add)): the fp32 -> operand-type cast in front of a GEMM, optionally with the "+ positional embedding" of
// the SAM attention inputs (sam/transformer.py:168-170,178-180) folded in. PDL: its launch overlaps the producer's tail.
template <bool BF16>
__global__ void cast16_kernel(const float4* __restrict__ x, const float4* __restrict__ add, uint2* __restrict__ y, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = x[i];
    if (add != nullptr) {
      const float4 a = add[i];
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    y[i] = make_uint2(pack2<BF16>(v.x, v.y), pack2<BF16>(v.z, v.w));
  }
}

// ------------------------------------------------------------------------------------------------
// K9: trilinear resampling of channels-last [B,Ti,Hi,Wi,C] -> [B,To,Ho,Wo,C]; one thread per (voxel, 8 channels).
// align_corners=1: src = dst*(in-1)/(out-1)   (dpt_block.py:231-236, dpt_head.py:81-83)
// align_corners=0: src = max((dst+0.5)*in/out - 0.5, 0)   (sparse_heads.py:645-647)
// ------------------------------------------------------------------------------------------------
struct Axis {
  int i0, i1;
  float w1;
};
L4P_DEVICE Axis axis_coord(int o, int in, int out, int align) {
  Axis a;
  float s;
  if (align) {
    s = out > 1 ? (float)o * ((float)(in - 1) / (float)(out - 1)) : 0.f;
  } else {
    s = ((float)o + 0.5f) * ((float)in / (float)out) - 0.5f;
    s = s < 0.f ? 0.f : s;
  }
  a.i0 = (int)s;
  if (a.i0 > in - 1) a.i0 = in - 1;
  a.i1 = a.i0 + 1 < in ? a.i0 + 1 : in - 1;
  a.w1 = s - (float)a.i0;
  return a;
}

// grid (ceil(Wo * C/8 / 256), ceil(Ho / kUpRows), B * To): the (b, t) coordinates and their source planes are
// block-uniform, a thread's (wo, channel group) comes from one 32-bit division and is reused for kUpRows output rows
// (round 2: one row per thread was instruction-bound - ncu 77 % issue utilisation at 0.25 of the HBM roofline - so the
// per-thread index math is amortised over four rows and the blend runs as packed f32x2 FMAs: 16 instead of 32 per corner
// quartet). All corner loads of a row are issued before its arithmetic.
constexpr int kUpRows = 4;
template <bool BF16>
__global__ void __launch_bounds__(256)
upsample_cl_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y,
                                   uint16_t* __restrict__ y_relu, int B, int Ti, int Hi, int Wi, int To, int Ho, int Wo,
                                   int C, int align, long long total) {
  const int cg = C >> 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (wo, g)
  if (idx >= Wo * cg) return;
  const int wo = idx / cg, g = idx - wo * cg;
  const int b = blockIdx.z / To, to = blockIdx.z - b * To;
  const Axis at = axis_coord(to, Ti, To, align);
  const Axis aw = axis_coord(wo, Wi, Wo, align);
  const bool two_t = at.w1 != 0.f;   // block-uniform: the DPT x2 upsampling keeps T
  const long long plane_t0 = ((long long)b * Ti + at.i0) * Hi, plane_t1 = ((long long)b * Ti + at.i1) * Hi;
  const long long o0 = (long long)aw.i0 * C + g * 8, o1 = (long long)aw.i1 * C + g * 8;
  const float ww0 = 1.f - aw.w1, ww1 = aw.w1, wt0 = 1.f - at.w1, wt1 = at.w1;
#pragma unroll
  for (int r = 0; r < kUpRows; ++r) {
    const int ho = blockIdx.y * kUpRows + r;
    if (ho >= Ho) break;
    const Axis ah = axis_coord(ho, Hi, Ho, align);
    const uint16_t* r00 = x + ((plane_t0 + ah.i0) * Wi) * (long long)C;
    const uint16_t* r01 = x + ((plane_t0 + ah.i1) * Wi) * (long long)C;
    uint4 raw[8];
    raw[0] = *reinterpret_cast<const uint4*>(r00 + o0); raw[1] = *reinterpret_cast<const uint4*>(r00 + o1);
    raw[2] = *reinterpret_cast<const uint4*>(r01 + o0); raw[3] = *reinterpret_cast<const uint4*>(r01 + o1);
    if (two_t) {
      const uint16_t* r10 = x + ((plane_t1 + ah.i0) * Wi) * (long long)C;
      const uint16_t* r11 = x + ((plane_t1 + ah.i1) * Wi) * (long long)C;
      raw[4] = *reinterpret_cast<const uint4*>(r10 + o0); raw[5] = *reinterpret_cast<const uint4*>(r10 + o1);
      raw[6] = *reinterpret_cast<const uint4*>(r11 + o0); raw[7] = *reinterpret_cast<const uint4*>(r11 + o1);
    }
    const float wh0 = 1.f - ah.w1, wh1 = ah.w1;
    const float wgt[8] = {wt0 * wh0 * ww0, wt0 * wh0 * ww1, wt0 * wh1 * ww0, wt0 * wh1 * ww1,
                          wt1 * wh0 * ww0, wt1 * wh0 * ww1, wt1 * wh1 * ww0, wt1 * wh1 * ww1};
    uint64_t acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = pk2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c >= 4 && !two_t) break;  // block-uniform
      const uint64_t w2 = pk2(wgt[c], wgt[c]);
      const uint32_t rr[4] = {raw[c].x, raw[c].y, raw[c].z, raw[c].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack2<BF16>(rr[i]);
        acc[i] = fma2(w2, pk2(f.x, f.y), acc[i]);
      }
    }
    float a[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) upk2(acc[i], a[2 * i], a[2 * i + 1]);
    const long long o = ((((long long)b * To + to) * Ho + ho) * (long long)Wo + wo) * C + g * 8;
    if (y != nullptr)
      *reinterpret_cast<uint4*>(y + o) = make_uint4(pack2<BF16>(a[0], a[1]), pack2<BF16>(a[2], a[3]),
                                                    pack2<BF16>(a[4], a[5]), pack2<BF16>(a[6], a[7]));
    if (y_relu != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], 0.f);
      *reinterpret_cast<uint4*>(y_relu + o) = make_uint4(pack2<BF16>(a[0], a[1]), pack2<BF16>(a[2], a[3]),
                                                         pack2<BF16>(a[4], a[5]), pack2<BF16>(a[6], a[7]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// N3: video preprocessing fused into one gather (l4p_dataset_mini.py:543-587 for the rgb key):
//   temporal mirror padding (:126-190: cat[x, flip(x)[1:]] until T >= T_out)  ->  spatial resize to (Hs, Ws)
//   (F.interpolate trilinear with T unchanged = bilinear, align_corners=False, :237-290)  ->  crop at (t0, i0, j0)
//   (:292-345)  ->  (x / 255 - mean) / std (:576-580).
// src: uint8 frames [T0, H0, W0, 3] (decoder layout) ; out: fp32 [3, T_out, Hc, Wc] (one clip of rgb_b3thw).
// One thread per output pixel (3 channels): the 4 taps x 3 bytes are gathered from L2, the three fp32 planes are
// written coalesced along x.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
preprocess_rgb_kernel(const uint8_t* __restrict__ src, float* __restrict__ out, int T0, int H0, int W0, int Hs, int Ws,
                      int t0, int i0, int j0, int To, int Hc, int Wc, float3 mean, float3 istd) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, t = blockIdx.z;
  if (x >= Wc) return;
  // mirror padding: the padded sequence is the period-(2 T0 - 2) reflection of the source
  int ts = t + t0;
  if (T0 > 1) {
    const int period = 2 * T0 - 2;
    ts %= period;
    if (ts >= T0) ts = period - ts;
  } else {
    ts = 0;
  }
  const Axis ay = axis_coord(y + i0, H0, Hs, 0);
  const Axis ax = axis_coord(x + j0, W0, Ws, 0);
  const uint8_t* f = src + (long long)ts * H0 * W0 * 3;
  const uint8_t* p00 = f + ((long long)ay.i0 * W0 + ax.i0) * 3;
  const uint8_t* p01 = f + ((long long)ay.i0 * W0 + ax.i1) * 3;
  const uint8_t* p10 = f + ((long long)ay.i1 * W0 + ax.i0) * 3;
  const uint8_t* p11 = f + ((long long)ay.i1 * W0 + ax.i1) * 3;
  const float w00 = (1.f - ay.w1) * (1.f - ax.w1), w01 = (1.f - ay.w1) * ax.w1, w10 = ay.w1 * (1.f - ax.w1), w11 = ay.w1 * ax.w1;
  const long long plane = (long long)To * Hc * Wc;
  const long long o = ((long long)t * Hc + y) * Wc + x;
  const float m[3] = {mean.x, mean.y, mean.z}, is[3] = {istd.x, istd.y, istd.z};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = (w00 * (float)p00[c] + w01 * (float)p01[c] + w10 * (float)p10[c] + w11 * (float)p11[c]) * (1.f / 255.f);
    out[c * plane + o] = (v - m[c]) * is[c];
  }
}

// ------------------------------------------------------------------------------------------------
// Strided 3x3x3 gather (pad 1): x [B,T,H,W,C] -> A [B*To*Ho*Wo, 27*C], K ordered (kt,kh,kw,c).
// Only used for the 1024->1024 stride-2 conv on the 8x16x16 grid (dpt_block.py:265-278): 14 MB.
// ------------------------------------------------------------------------------------------------
__global__ void im2col3_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int T, int H, int W, int C,
                               int sT, int sH, int sW, int To, int Ho, int Wo, long long total) {
  const int cg = C / 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    long long v = idx / cg;
    const int tap = (int)(v % 27); v /= 27;
    const int wo = (int)(v % Wo); v /= Wo;
    const int ho = (int)(v % Ho); v /= Ho;
    const int to = (int)(v % To); v /= To;
    const long long b = v;
    const int ti = to * sT + tap / 9 - 1;
    const int hi = ho * sH + (tap / 3) % 3 - 1;
    const int wi = wo * sW + tap % 3 - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (ti >= 0 && ti < T && hi >= 0 && hi < H && wi >= 0 && wi < W)
      val = x[(((b * T + ti) * H + hi) * (long long)W + wi) * cg + g];
    out[idx] = val;
  }
}

static unsigned grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)host_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace l4p

using namespace l4p;

extern "C" int l4p_patchify(const float* rgb, void* out16, int B, int C, int T, int H, int W, int pt, int ph, int pw,
                            int bf16, void* stream) {
  L4P_REQUIRE(rgb && out16, L4P_ERR_ARG, "l4p_patchify: null pointer");
  L4P_REQUIRE(B > 0 && C > 0 && pt > 0 && ph > 0 && pw > 0 && T % pt == 0 && H % ph == 0 && W % pw == 0, L4P_ERR_SHAPE,
              "l4p_patchify: [%d,%d,%d,%d,%d] not divisible by tubelet (%d,%d,%d)", B, C, T, H, W, pt, ph, pw);
  const long long total = (long long)B * C * T * H * W;
  const unsigned grid = grid_for(total, 256);
  if (bf16)
    patchify_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(rgb, (uint16_t*)out16, B, C, T, H, W, pt, ph, pw, total);
  else
    patchify_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(rgb, (uint16_t*)out16, B, C, T, H, W, pt, ph, pw, total);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}

extern "C" int l4p_cast16_add(const float* x, const float* add, void* y16, int64_t n, int bf16, void* stream) {
  L4P_REQUIRE(x && y16, L4P_ERR_ARG, "l4p_cast16: null pointer");
  L4P_REQUIRE(n >= 0 && n % 4 == 0, L4P_ERR_SHAPE, "l4p_cast16: n=%lld must be a multiple of 4", (long long)n);
  if (n == 0) return L4P_OK;
  const unsigned grid = grid_for(n / 4, 256);
  if (bf16)
    L4P_CHECK_CUDA(launch_pdl(cast16_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, (const float4*)add, (uint2*)y16, (long long)(n / 4)));
  else
    L4P_CHECK_CUDA(launch_pdl(cast16_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float4*)x, (const float4*)add, (uint2*)y16, (long long)(n / 4)));
  return L4P_OK;
}

extern "C" int l4p_cast16(const float* x, void* y16, int64_t n, int bf16, void* stream) {
  return l4p_cast16_add(x, nullptr, y16, n, bf16, stream);
}

extern "C" int l4p_upsample3d(const void* x16, void* y16, void* y16_relu, int B, int Ti, int Hi, int Wi, int To, int Ho,
                              int Wo, int C, int align_corners, int bf16, void* stream) {
  L4P_REQUIRE(x16 && (y16 || y16_relu), L4P_ERR_ARG, "l4p_upsample3d: null pointer");
  L4P_REQUIRE(B > 0 && Ti > 0 && Hi > 0 && Wi > 0 && To > 0 && Ho > 0 && Wo > 0 && C > 0 && C % 8 == 0, L4P_ERR_SHAPE,
              "l4p_upsample3d: bad shape (C=%d must be a multiple of 8)", C);
  const long long total = (long long)B * To * Ho * Wo * (C / 8);
  L4P_REQUIRE((long long)Wo * (C / 8) < (1ll << 31) && Ho <= 65535 && (long long)B * To <= 65535, L4P_ERR_SHAPE,
              "l4p_upsample3d: grid too large");
  const dim3 grid((unsigned)(((long long)Wo * (C / 8) + 255) / 256), (unsigned)((Ho + kUpRows - 1) / kUpRows), (unsigned)(B * To));
  if (bf16)
    upsample_cl_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)x16, (uint16_t*)y16,
                                                                       (uint16_t*)y16_relu, B, Ti, Hi, Wi, To, Ho, Wo, C,
                                                                       align_corners, total);
  else
    upsample_cl_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)x16, (uint16_t*)y16,
                                                                        (uint16_t*)y16_relu, B, Ti, Hi, Wi, To, Ho, Wo, C,
                                                                        align_corners, total);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}

extern "C" int l4p_im2col3(const void* x16, void* out16, int B, int T, int H, int W, int C, int sT, int sH, int sW,
                           void* stream) {
  L4P_REQUIRE(x16 && out16, L4P_ERR_ARG, "l4p_im2col3: null pointer");
  L4P_REQUIRE(B > 0 && T > 0 && H > 0 && W > 0 && C % 8 == 0 && sT > 0 && sH > 0 && sW > 0, L4P_ERR_SHAPE,
              "l4p_im2col3: bad shape");
  const int To = (T + 2 - 3) / sT + 1, Ho = (H + 2 - 3) / sH + 1, Wo = (W + 2 - 3) / sW + 1;
  const long long total = (long long)B * To * Ho * Wo * 27 * (C / 8);
  im2col3_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)x16, (uint4*)out16, B, T, H, W, C,
                                                                        sT, sH, sW, To, Ho, Wo, total);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}

extern "C" int l4p_preprocess_rgb(const void* frames_u8, float* out, int T0, int H0, int W0, int Hs, int Ws, int t0, int i0,
                                  int j0, int To, int Hc, int Wc, const float* mean3, const float* std3, void* stream) {
  L4P_REQUIRE(frames_u8 && out && mean3 && std3, L4P_ERR_ARG, "l4p_preprocess_rgb: null pointer");
  L4P_REQUIRE(T0 > 0 && H0 > 0 && W0 > 0 && Hs > 0 && Ws > 0 && To > 0 && Hc > 0 && Wc > 0, L4P_ERR_SHAPE,
              "l4p_preprocess_rgb: empty shape");
  L4P_REQUIRE(t0 >= 0 && i0 >= 0 && j0 >= 0 && i0 + Hc <= Hs && j0 + Wc <= Ws, L4P_ERR_SHAPE,
              "l4p_preprocess_rgb: crop (%d,%d)+(%d,%d) outside the resized frame %dx%d", i0, j0, Hc, Wc, Hs, Ws);
  L4P_REQUIRE(Hc <= 65535 && To <= 65535, L4P_ERR_SHAPE, "l4p_preprocess_rgb: grid too large");
  const dim3 grid((unsigned)((Wc + 255) / 256), (unsigned)Hc, (unsigned)To);
  preprocess_rgb_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)frames_u8, out, T0, H0, W0, Hs, Ws, t0, i0, j0,
                                                                 To, Hc, Wc, make_float3(mean3[0], mean3[1], mean3[2]),
                                                                 make_float3(1.f / std3[0], 1.f / std3[1], 1.f / std3[2]));
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}
