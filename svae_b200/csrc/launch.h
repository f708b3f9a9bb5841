// launch.h -- per-dimension launcher declarations (one definition per lds_estep_inst.cu object).
#pragma once
#include <cuda_runtime.h>
#include "lds_estep.cuh"
#include "lds_estep_bwd.cuh"

namespace svae {
#define SVAE_DECL_LDS(n) cudaError_t lds_fwd_launch_n##n(const LdsFwdParams&, int cpl, int lps, bool inhomog, cudaStream_t);
SVAE_DECL_LDS(1) SVAE_DECL_LDS(2) SVAE_DECL_LDS(3) SVAE_DECL_LDS(4) SVAE_DECL_LDS(5) SVAE_DECL_LDS(6)
SVAE_DECL_LDS(7) SVAE_DECL_LDS(8) SVAE_DECL_LDS(9) SVAE_DECL_LDS(10) SVAE_DECL_LDS(11) SVAE_DECL_LDS(12)
SVAE_DECL_LDS(13) SVAE_DECL_LDS(14) SVAE_DECL_LDS(15) SVAE_DECL_LDS(16)
#undef SVAE_DECL_LDS
#define SVAE_DECL_LDSB(n) cudaError_t lds_bwd_launch_n##n(const LdsBwdParams&, int lps, cudaStream_t);
SVAE_DECL_LDSB(1) SVAE_DECL_LDSB(2) SVAE_DECL_LDSB(3) SVAE_DECL_LDSB(4) SVAE_DECL_LDSB(5) SVAE_DECL_LDSB(6)
SVAE_DECL_LDSB(7) SVAE_DECL_LDSB(8) SVAE_DECL_LDSB(9) SVAE_DECL_LDSB(10) SVAE_DECL_LDSB(11) SVAE_DECL_LDSB(12)
SVAE_DECL_LDSB(13) SVAE_DECL_LDSB(14) SVAE_DECL_LDSB(15) SVAE_DECL_LDSB(16)
#undef SVAE_DECL_LDSB
}  // namespace svae
