// Micro-benchmark: sustained tcgen05.mma issue rate of one CTA per SM as a function of N, for SS (A and B in shared
// memory) and TS (A in TMEM) operands, M = 128, K = 16 per instruction, kind::f16. Operands are never loaded
// (zero-filled smem / uninitialised TMEM): only the tensor-pipe + operand-fetch time is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu && ./umma_bench
#include <cstdio>
#include <cstdlib>
#include "../../l4p_b200/csrc/common.cuh"

using namespace l4p;

template <int NACC, int WARPS>
__global__ void __launch_bounds__(128, 1) bench_kernel(int N, int ts_mode, int iters, int nacc, long long* out) {
  const int swz64 = 0;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32768;  // A: 128 x 64 (16 KiB), B: up to 256 x 64 (32 KiB)
  for (int i = threadIdx.x; i < (32768 + 65536) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), WARPS); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  // WARPS issuing threads (lane 0 of warps 0..WARPS-1), each with its own accumulator range: is the limit per thread or per SM?
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < WARPS) {
    const uint32_t wtm = tm + (threadIdx.x >> 5) * 64;
    const uint32_t idesc = umma_idesc_f16(false, 128, (uint32_t)N);
    const uint32_t hi = swz64 ? umma_desc_hi(64, 4) : umma_desc_hi(128, 2);
    const uint32_t a_lo = umma_desc_lo(sA), b_lo = umma_desc_lo(sB);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // nacc independent accumulators used round-robin (dependent-chain latency vs issue rate); N * nacc <= 256 columns
        const uint32_t d = wtm + (uint32_t)((k % NACC) * N);
        if (ts_mode) umma_ts(d, tm + 256 + k * 8, umma_desc_make(b_lo + 2 * k, hi), idesc, 1u);
        else umma_ss(d, umma_desc_make(a_lo + 2 * k, hi), umma_desc_make(b_lo + 2 * k, hi), idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);  // count = WARPS
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

namespace l4p {
int host_set_error(int code, const char*, ...) { return code; }
int host_check_cuda(cudaError_t e, const char*) { return e == cudaSuccess ? 0 : -1; }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  const int iters = 2000;
  const int grid = 148;
  auto run = [&](auto kern, const char* name, int N, int ts, int warps) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    kern<<<grid, 128, 100 * 1024>>>(N, ts, 10, 0, d_out);
    kern<<<grid, 128, 100 * 1024>>>(N, ts, iters, 0, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%s %s N=%3d: %7.1f cycles / UMMA per issuing thread, %7.1f per SM (nominal %5.1f) %s\n", name, ts ? "TS" : "SS", N,
           (double)cyc / (4.0 * iters), (double)cyc / (4.0 * iters * warps), 128.0 * N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {16, 32, 64}) {
      run(bench_kernel<1, 1>, "1 thread, 1 accumulator ", N, ts, 1);
      run(bench_kernel<2, 1>, "1 thread, 2 accumulators", N, ts, 1);
      run(bench_kernel<4, 1>, "1 thread, 4 accumulators", N, ts, 1);
      run(bench_kernel<1, 2>, "2 threads, 1 acc each   ", N, ts, 2);
      run(bench_kernel<1, 4>, "4 threads, 1 acc each   ", N, ts, 4);
    }
  return 0;
}
