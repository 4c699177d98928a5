// Explicit instantiations of the GEMM kernels for one group of epilogue configurations (see gemm_kernel.cuh).
#include "gemm_kernel.cuh"

namespace l4p {

const GemmKernelSet* gemm_instances_d(int* n) {
  static const GemmKernelSet sets[] = {
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_HEAD1X1, L4P_ACT_RELU, 0)),
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_HYPER, L4P_ACT_GELU, 0)),
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_ROWMAJOR, 0, EPI_GENERIC)),
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_CONVT, 0, EPI_GENERIC)),
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_HEAD1X1, 0, EPI_GENERIC)),
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_HYPER, 0, EPI_GENERIC)),
  };
  *n = (int)(sizeof(sets) / sizeof(sets[0]));
  return sets;
}

void (*gemm_splitk_finalize(bool bf16))(const GemmKParams) {
  return bf16 ? splitk_finalize_kernel<true> : splitk_finalize_kernel<false>;
}

}  // namespace l4p
