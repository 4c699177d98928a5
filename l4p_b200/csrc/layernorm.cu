// K2: LayerNorm over the channel axis, one warp per row, fp32 statistics (two-pass in registers),
// 16-bit (GEMM operand) and/or fp32 output. HBM-bound: one read of x, one write of y.
// Reference call sites: modeling_finetune.py:247-248, l4p_videomae.py:115, sam/transformer.py:139-149.
#include "common.cuh"
#include "../../include/l4p_b200.h"

namespace l4p {

constexpr int kLnMaxVec = 12;  // float4 per lane: cols <= 12*4*32 = 1536

// PRE: fetch gamma / beta before the programmatic-dependency wait (the launch-latency-bound small-row case: +96 registers
// are free there); the large-row, bandwidth-bound case keeps the registers for occupancy and reads them at the end.
template <bool BF16, bool PRE>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 uint16_t* __restrict__ y16, float* __restrict__ y32, long long rows, int cols, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const int nvec = cols >> 2;  // float4 per row
  // gamma / beta are weights: fetched BEFORE the programmatic-dependency wait, so their L2 latency overlaps the
  // predecessor's tail instead of sitting between the two reductions and the store
  float4 gam[PRE ? kLnMaxVec : 1], bet[PRE ? kLnMaxVec : 1];
  if constexpr (PRE) {
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int j = lane + i * 32;
      if (j < nvec) {
        gam[i] = reinterpret_cast<const float4*>(gamma)[j];
        bet[i] = reinterpret_cast<const float4*>(beta)[j];
      }
    }
  }
  pdl_launch_dependents();
  pdl_wait();
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
  float4 v[kLnMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      v[i] = xr[j];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      const float4 g = PRE ? gam[PRE ? i : 0] : reinterpret_cast<const float4*>(gamma)[j];
      const float4 b = PRE ? bet[PRE ? i : 0] : reinterpret_cast<const float4*>(beta)[j];
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (y32 != nullptr) reinterpret_cast<float4*>(y32 + row * cols)[j] = o;
      if (y16 != nullptr)
        reinterpret_cast<uint2*>(y16 + row * cols)[j] = make_uint2(pack2<BF16>(o.x, o.y), pack2<BF16>(o.z, o.w));
    }
  }
}

}  // namespace l4p

using namespace l4p;

extern "C" int l4p_layernorm(const float* x, const float* gamma, const float* beta, void* y16, float* y32,
                             int64_t rows, int cols, float eps, int bf16, void* stream) {
  L4P_REQUIRE(x && gamma && beta && (y16 || y32), L4P_ERR_ARG, "l4p_layernorm: null pointer");
  L4P_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0 && cols <= kLnMaxVec * 128, L4P_ERR_SHAPE,
              "l4p_layernorm: cols=%d (multiple of 4, <= %d)", cols, kLnMaxVec * 128);
  if (rows == 0) return L4P_OK;
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  // small row counts (the encoder: 2048 rows = one wave) are launch-latency-bound: PDL; large ones launch normally so
  // that an early-scheduled successor cannot take SM slots from this kernel's later waves
  const long long llrows = rows;
  if (grid <= 2048u) {
    if (bf16)
      L4P_CHECK_CUDA(launch_pdl(layernorm_kernel<true, true>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, x, gamma, beta, (uint16_t*)y16, y32, llrows, cols, eps));
    else
      L4P_CHECK_CUDA(launch_pdl(layernorm_kernel<false, true>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, x, gamma, beta, (uint16_t*)y16, y32, llrows, cols, eps));
    return L4P_OK;
  }
  if (bf16)
    layernorm_kernel<true, false><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(x, gamma, beta, (uint16_t*)y16, y32, rows, cols, eps);
  else
    layernorm_kernel<false, false><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(x, gamma, beta, (uint16_t*)y16, y32, rows, cols, eps);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}
