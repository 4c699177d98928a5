// Micro-benchmark 3 (round 2): what ONE producer thread can pull through TMA per SM. A ring of S smem stages, one mbarrier each;
// the thread waits for the stage's previous load, arms expect_tx and issues `per_stage` cp.async.bulk.tensor.2d loads
// (box = 64 x rows 16-bit elements, 128-byte swizzle) from one or two tensor maps, exactly as the GEMM producer does.
// Reports cycles per ring iteration and bytes / clk / SM for a grid of 1 and of 148 CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_issue_bench tma_issue_bench.cu -lcuda && ./tma_issue_bench
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "../../l4p_b200/csrc/common.cuh"

using namespace l4p;

struct Case { const char* name; int rows_a, rows_b; int two_maps; int stages; };

__global__ void __launch_bounds__(128, 1)
bench_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int rows_a, int rows_b, int stages,
             int iters, int m_rows, long long* out, long long* trace) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[16];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)rows_a * 128u, b_bytes = (uint32_t)rows_b * 128u;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bar[s]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int stage = 0;
    uint32_t phase = 0;
    const int row0 = (int)((blockIdx.x * 256) % (unsigned)m_rows);
    long long t0 = 0;
    for (int i = 0; i < iters + stages; ++i) {
      const bool tr = trace != nullptr && blockIdx.x == 0 && i < 48;
      const long long c0 = clock64();
      if (i >= stages) mbar_wait(smem_u32(&bar[stage]), phase ^ 1u);   // the load issued `stages` iterations ago has landed
      if (i == stages) t0 = clock64();
      const long long c1 = clock64();
      const uint32_t sa = base + (uint32_t)stage * stage_bytes;
      mbar_expect_tx(smem_u32(&bar[stage]), a_bytes + b_bytes);
      const long long c2 = clock64();
      const int kc = (i * 64) % 4096;
      tma_load_2d(sa, &tmA, smem_u32(&bar[stage]), kc, row0);
      const long long c3 = clock64();
      if (rows_b > 0) tma_load_2d(sa + a_bytes, &tmB, smem_u32(&bar[stage]), kc, row0);
      const long long c4 = clock64();
      if (tr) { trace[i * 5] = c0; trace[i * 5 + 1] = c1; trace[i * 5 + 2] = c2; trace[i * 5 + 3] = c3; trace[i * 5 + 4] = c4; }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    const long long t1 = clock64();
    // drain
    for (int s = 0; s < stages; ++s) {
      mbar_wait(smem_u32(&bar[stage]), phase ^ 1u);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    out[blockIdx.x] = t1 - t0;
  }
}


// Loop-structure variants for ONE load per iteration (the A warp of the GEMM): where do the ~250 cycles go?
//   mode 0: wait, expect_tx, TMA (the plain loop)          mode 1: the NEXT stage's try_wait is issued right after this stage's TMA
//   mode 2: no waits at all (lower bound: expect_tx + TMA)  mode 3: no waits, no expect_tx (TMA issue alone; barriers never complete)
__global__ void __launch_bounds__(128, 1)
variant_kernel(const __grid_constant__ CUtensorMap tmA, int rows_a, int stages, int iters, int m_rows, int mode, int loads, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[16];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)rows_a * 128u;
  const uint32_t stage_bytes = ((uint32_t)loads * a_bytes + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bar[s]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int stage = 0;
    uint32_t phase = 0;
    const int row0 = (int)((blockIdx.x * 256) % (unsigned)m_rows);
    const long long t0 = clock64();
    bool ready = true;   // first pass: fresh barriers
    for (int i = 0; i < iters; ++i) {
      const uint32_t b = smem_u32(&bar[stage]);
      if (mode == 0) {
        mbar_wait(b, phase ^ 1u);
      } else if (mode == 1) {
        if (!ready) mbar_wait(b, phase ^ 1u);
      }
      const uint32_t sa = base + (uint32_t)stage * stage_bytes;
      if (mode != 3) mbar_expect_tx(b, (uint32_t)loads * a_bytes);
      const int kc = (i * 64) % 4096;
      for (int l = 0; l < loads; ++l) tma_load_2d(sa + (uint32_t)l * a_bytes, &tmA, b, kc, row0 + l * rows_a);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
      if (mode == 1) ready = mbar_try_wait(smem_u32(&bar[stage]), phase ^ 1u);
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  // no drain: the CTA exits with loads in flight only in modes 2 / 3 (harmless for a timing kernel: wait a while instead)
  if (threadIdx.x == 0) { const long long t = clock64(); while (clock64() - t < 20000) { } }
}

static void make_map(CUtensorMap* m, void* ptr, int K, int M, int box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
}

int main() {
  cudaFree(0);
  const int K = 4096, M = 148 * 256;
  void *a, *b;
  cudaMalloc(&a, (size_t)K * M * 2);
  cudaMalloc(&b, (size_t)K * M * 2);
  cudaMemset(a, 0, (size_t)K * M * 2);
  cudaMemset(b, 0, (size_t)K * M * 2);
  long long *out, *trace;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaMalloc(&trace, 48 * 5 * sizeof(long long));
  const Case cases[] = {
      {"A 128 rows + B 128 rows, two maps (pair 256x256 tile)", 128, 128, 1, 5},
      {"A 128 rows + B 128 rows, ONE map for both", 128, 128, 0, 5},
      {"A 128 rows + B 88 rows, two maps (pair N=176)", 128, 88, 1, 6},
      {"A 128 rows + B 32 rows, two maps (pair N=64)", 128, 32, 1, 8},
      {"A 64 rows + B 64 rows, two maps", 64, 64, 1, 8},
      {"A 128 rows only (one load per iteration)", 128, 0, 1, 8},
      {"A 256 rows only (one 32 KiB load per iteration)", 256, 0, 1, 5},
      {"A 64 rows only", 64, 0, 1, 8},
      {"A 32 rows only", 32, 0, 1, 8},
  };
  const int iters = 2000;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (const Case& c : cases) {
    CUtensorMap tmA, tmB;
    make_map(&tmA, a, K, M, c.rows_a);
    make_map(&tmB, c.two_maps ? b : a, K, M, c.rows_b > 0 ? c.rows_b : 8);
    if (!c.two_maps) tmB = tmA;
    for (int grid : {1, 148}) {
      for (int rep = 0; rep < 2; ++rep)
        bench_kernel<<<grid, 128, 200 * 1024>>>(tmA, c.two_maps ? tmB : tmA, c.rows_a, c.rows_b, c.stages, iters, M, out, trace);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double mean = 0, mx = 0;
      for (int i = 0; i < grid; ++i) { mean += (double)h[i]; if ((double)h[i] > mx) mx = (double)h[i]; }
      mean /= grid;
      const double bytes = (double)(c.rows_a + c.rows_b) * 128.0;
      if (grid == 1) {
        long long t[48 * 5];
        cudaMemcpy(t, trace, sizeof(t), cudaMemcpyDeviceToHost);
        printf("  trace (iteration: start-to-start, wait, expect_tx, tma A, tma B):");
        for (int i = 0; i < 20; ++i)
          printf(" [%d: %lld | %lld %lld %lld %lld]", i, i ? t[i * 5] - t[(i - 1) * 5] : 0ll, t[i * 5 + 1] - t[i * 5], t[i * 5 + 2] - t[i * 5 + 1],
                 t[i * 5 + 3] - t[i * 5 + 2], t[i * 5 + 4] - t[i * 5 + 3]);
        printf("\n");
      }
      printf("%-58s stages %d grid %3d: %7.1f cycles / iteration (max CTA %7.1f) = %5.1f B/clk/SM\n", c.name, c.stages, grid, mean / iters,
             mx / iters, bytes / (mean / iters));
    }
  }
  cudaFuncSetAttribute(variant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* mname[] = {"wait, expect_tx, TMA", "next stage's try_wait issued after this TMA", "no waits (expect_tx + TMA)", "TMA issue alone"};
  for (int loads : {1, 3})
    for (int mode = 0; mode < 4; ++mode) {
      CUtensorMap tmA;
      make_map(&tmA, a, K, M, 64);
      for (int rep = 0; rep < 2; ++rep) variant_kernel<<<1, 128, 200 * 1024>>>(tmA, 64, 8, iters, M, mode, loads, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      long long h;
      cudaMemcpy(&h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%d load(s) of 64 rows per iteration, %-45s: %7.1f cycles / iteration\n", loads, mname[mode], (double)h / iters);
    }
  return 0;
}
