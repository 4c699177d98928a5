// Explicit instantiations of the GEMM kernels for one group of epilogue configurations (see gemm_kernel.cuh).
#include "gemm_kernel.cuh"

namespace l4p {

const GemmKernelSet* gemm_instances_c(int* n) {
  static const GemmKernelSet sets[] = {
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_ROWMAJOR, L4P_ACT_NONE, EPI_RES16 | EPI_OUT16 | EPI_OUT16R)),
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_ROWMAJOR, L4P_ACT_NONE, EPI_RES16 | EPI_OUT16)),
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_ROWMAJOR, L4P_ACT_NONE, EPI_OUT16 | EPI_OUT16R)),
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_CONVT, L4P_ACT_NONE, 0)),
      L4P_GEMM_KERNEL_SET(epi_make(kStoreSplitK, L4P_ACT_NONE, 0)),
      L4P_GEMM_KERNEL_SET(epi_make(L4P_STORE_ROWMAJOR, L4P_ACT_NONE, EPI_RES16 | EPI_OUT16 | EPI_WIDE3)),
  };
  *n = (int)(sizeof(sets) / sizeof(sets[0]));
  return sets;
}

}  // namespace l4p
