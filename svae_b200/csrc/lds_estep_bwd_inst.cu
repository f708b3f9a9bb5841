// One translation unit per latent dimension for the VJP kernel (compiled with -DSVAE_N=<n>).
#include "lds_estep_bwd.cuh"
#include "launch.h"

#ifndef SVAE_N
#error "compile with -DSVAE_N=<latent dim>"
#endif

namespace svae {

template <int N, int LPS, bool PER_T>
static cudaError_t launch_bwd_t(const LdsBwdParams& p, cudaStream_t stream) {
  constexpr int SPW = 32 / LPS;
  const long long tiles = (p.B + SPW - 1) / SPW;
  const size_t smem = sizeof(double) * (size_t)SPW * lds_bwd_tile_doubles(N, p.S);
  auto kern = lds_estep_bwd_kernel<N, LPS, PER_T>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)tiles, 32, smem, stream>>>(p);
  return cudaGetLastError();
}

template <int N, int LPS>
static cudaError_t launch_bwd(const LdsBwdParams& p, cudaStream_t stream) {
  const bool per_t = p.g_Exx_t != nullptr || p.g_Exxn_t != nullptr;
  return per_t ? launch_bwd_t<N, LPS, true>(p, stream) : launch_bwd_t<N, LPS, false>(p, stream);
}

#define SVAE_CAT_(a, b) a##b
#define SVAE_CAT(a, b) SVAE_CAT_(a, b)

cudaError_t SVAE_CAT(lds_bwd_launch_n, SVAE_N)(const LdsBwdParams& p, int lps, cudaStream_t stream) {
  constexpr int N = SVAE_N;
  if (lps == 16) {
    if constexpr (N + 1 <= 16) {
      return launch_bwd<N, 16>(p, stream);
    } else {
      return cudaErrorInvalidValue;
    }
  }
  return launch_bwd<N, 32>(p, stream);
}

}  // namespace svae
