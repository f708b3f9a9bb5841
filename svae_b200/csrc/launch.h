// launch.h -- per-dimension launcher declarations (one definition per lds_estep_inst.cu object).
#pragma once
#include <cuda_runtime.h>
#include "lds_estep.cuh"

namespace svae {
#define SVAE_DECL_LDS(n) cudaError_t lds_fwd_launch_n##n(const LdsFwdParams&, int lps, bool inhomog, cudaStream_t);
SVAE_DECL_LDS(1) SVAE_DECL_LDS(2) SVAE_DECL_LDS(3) SVAE_DECL_LDS(4) SVAE_DECL_LDS(5) SVAE_DECL_LDS(6)
SVAE_DECL_LDS(7) SVAE_DECL_LDS(8) SVAE_DECL_LDS(9) SVAE_DECL_LDS(10) SVAE_DECL_LDS(11) SVAE_DECL_LDS(12)
SVAE_DECL_LDS(13) SVAE_DECL_LDS(14) SVAE_DECL_LDS(15) SVAE_DECL_LDS(16)
#undef SVAE_DECL_LDS
}  // namespace svae
