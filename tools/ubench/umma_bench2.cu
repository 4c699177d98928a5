// Micro-benchmark 2 (round 2): cost of the EXACT tcgen05.mma shapes / descriptor patterns of csrc/attention.cu, one issuing
// thread per SM, operands resident (zero-filled) in shared memory / TMEM, accumulation chains as in the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench2 umma_bench2.cu && ./umma_bench2
// Cases: S-shaped UMMAs (M128 N128 K16, SS) with 64-byte-swizzled chunks (the kernel's Q/K layout: three 32-element chunks)
// vs 128-byte-swizzled chunks; PV-shaped UMMAs (M128 N96 K16) SS and TS; N=256; and the kernel's whole per-key-block sequence.
#include <cstdio>
#include <cstdlib>
#include "../../l4p_b200/csrc/common.cuh"

using namespace l4p;

enum Case { S_SWZ64 = 0, S_SWZ128 = 1, PV_SS = 2, PV_TS = 3, SS_N256 = 4, BLOCK_SEQ64 = 5, BLOCK_SEQ128 = 6, S_SWZ64_N64 = 7,
            S_SWZ128_K128 = 8, NCASES = 9 };

__global__ void __launch_bounds__(128, 1) bench_kernel(int which, int iters, long long* out, int* count) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // Q tiles 2 x 24 KiB | K 24 KiB | V 24 KiB | P 32 KiB  (as in the kernel, one ring slot each)
  const uint32_t sQ = base, sK = sQ + 2 * 24576, sV = sK + 24576, sP = sV + 24576;
  for (int i = threadIdx.x; i < (4 * 24576 + 32768) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc_s = umma_idesc_f16(false, 128, 128), idesc_o = umma_idesc_f16(false, 128, 96);
    const uint32_t idesc_256 = umma_idesc_f16(false, 128, 256), idesc_s64 = umma_idesc_f16(false, 128, 64);
    constexpr uint32_t hi64 = umma_desc_hi(64, 4), hi128 = umma_desc_hi(128, 2);
    const uint32_t q_lo = umma_desc_lo(sQ), k_lo = umma_desc_lo(sK), p_lo = umma_desc_lo(sP), v_lo = umma_desc_lo(sV);
    int n = 0;
    auto s64 = [&](int t, uint32_t idesc) {   // the kernel's issue_s: 6 K-steps over three 64B-swizzled chunks of 128 rows
      const uint32_t d = tm + (t ? 128u : 0u);
#pragma unroll
      for (int kk = 0; kk < 6; ++kk) {
        const uint32_t off = ((uint32_t)(kk >> 1) * (128 * 64) + (uint32_t)(kk & 1) * 32) >> 4;
        umma_ss(d, umma_desc_make(q_lo + (uint32_t)t * (24576 >> 4) + off, hi64), umma_desc_make(k_lo + off, hi64), idesc, kk != 0);
      }
      n += 6;
    };
    auto s128 = [&](int t, int ksteps) {      // same math with 128B-swizzled chunks of 64 elements (4 K-steps per chunk)
      const uint32_t d = tm + (t ? 128u : 0u);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (kk < ksteps) {
          const uint32_t off = ((uint32_t)(kk >> 2) * (128 * 128) + (uint32_t)(kk & 3) * 32) >> 4;
          umma_ss(d, umma_desc_make(q_lo + off, hi128), umma_desc_make(k_lo + off, hi128), idesc_s, kk != 0);
        }
      }
      n += ksteps;
    };
    auto pv = [&](int t, bool ts) {           // the kernel's issue_pv: 8 K-steps, V^T two 128B-swizzled chunks of 96 rows
      const uint32_t d = tm + (t ? 352u : 256u);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t o = ((uint32_t)(kk & 3) * 32) >> 4;
        const uint64_t vdesc = umma_desc_make(v_lo + (uint32_t)(kk >> 2) * ((96 * 128) >> 4) + o, hi128);
        if (ts) umma_ts(d, tm + 448 + (uint32_t)kk * 8u, vdesc, idesc_o, kk != 0);
        else umma_ss(d, umma_desc_make(p_lo + (uint32_t)(kk >> 2) * ((128 * 128) >> 4) + o, hi128), vdesc, idesc_o, kk != 0);
      }
      n += 8;
    };
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      switch (which) {
        case S_SWZ64: s64(it & 1, idesc_s); break;
        case S_SWZ64_N64: s64(it & 1, idesc_s64); break;
        case S_SWZ128: s128(it & 1, 6); break;
        case S_SWZ128_K128: s128(it & 1, 8); break;
        case PV_SS: pv(it & 1, false); break;
        case PV_TS: pv(it & 1, true); break;
        case SS_N256: {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tm + (it & 1) * 256u, umma_desc_make(q_lo + 2 * kk, hi128), umma_desc_make(k_lo + 2 * kk, hi128), idesc_256, kk != 0);
          n += 4;
          break;
        }
        case BLOCK_SEQ64: s64(0, idesc_s); pv(0, true); s64(1, idesc_s); pv(1, false); break;
        case BLOCK_SEQ128: s128(0, 6); pv(0, true); s128(1, 6); pv(1, false); break;
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; count[0] = n; }
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
  long long* d_out; int* d_cnt;
  cudaMalloc(&d_out, 8); cudaMalloc(&d_cnt, 4);
  const char* names[NCASES] = {"S  SS swz64  M128 N128 (kernel Q/K layout, 6 K-steps)", "S  SS swz128 M128 N128 (6 K-steps)",
                               "PV SS swz128 M128 N96  (8 K-steps)", "PV TS        M128 N96  (8 K-steps)",
                               "   SS swz128 M128 N256 (4 K-steps)", "key block: S0 PV0(TS) S1 PV1(SS), swz64 S",
                               "key block: S0 PV0(TS) S1 PV1(SS), swz128 S", "S  SS swz64  M128 N64  (6 K-steps)",
                               "S  SS swz128 M128 N128 (8 K-steps, d=128)"};
  const double nominal[NCASES] = {64, 64, 48, 48, 128, 0, 0, 32, 64};
  const int smem = 4 * 24576 + 32768 + 1024;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148})
    for (int c = 0; c < NCASES; ++c) {
      bench_kernel<<<grid, 128, smem>>>(c, 10, d_out, d_cnt);
      bench_kernel<<<grid, 128, smem>>>(c, 1000, d_out, d_cnt);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0; int n = 0;
      cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(&n, d_cnt, 4, cudaMemcpyDeviceToHost);
      printf("grid %3d | %-52s: %7.1f cycles / UMMA (nominal %5.1f), %8.1f cycles / iteration %s\n", grid, names[c],
             (double)cyc / n, nominal[c], (double)cyc / 1000.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
