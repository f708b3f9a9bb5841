// One translation unit per latent dimension: compiled with -DSVAE_N=<n> for n = 1..SVAE_LDS_MAX_N so
// the fully-unrolled kernels build in parallel.  Exposes svae::lds_fwd_launch_n<n>().
#include "lds_estep.cuh"
#include "launch.h"

#ifndef SVAE_N
#error "compile with -DSVAE_N=<latent dim>"
#endif

namespace svae {

template <int N, int LPS, bool INHOMOG>
static cudaError_t launch_one(const LdsFwdParams& p, cudaStream_t stream) {
  constexpr int SPW = 32 / LPS;
  const int warps_per_cta = 1;
  const long long tiles = (p.B + SPW - 1) / SPW;
  const long long grid = (tiles + warps_per_cta - 1) / warps_per_cta;
  const size_t smem = sizeof(double) *
      (3 * N * lds_np(N) + (size_t)warps_per_cta * SPW * lds_tile_doubles(N, p.S, INHOMOG));
  auto kern = lds_estep_fwd_kernel<N, LPS, INHOMOG>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)grid, warps_per_cta * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

#define SVAE_CAT_(a, b) a##b
#define SVAE_CAT(a, b) SVAE_CAT_(a, b)

cudaError_t SVAE_CAT(lds_fwd_launch_n, SVAE_N)(const LdsFwdParams& p, int lps, bool inhomog, cudaStream_t stream) {
  constexpr int N = SVAE_N;
  if (lps == 16) {
    if constexpr (N + 1 <= 16) {
      return inhomog ? launch_one<N, 16, true>(p, stream) : launch_one<N, 16, false>(p, stream);
    } else {
      return cudaErrorInvalidValue;
    }
  }
  return inhomog ? launch_one<N, 32, true>(p, stream) : launch_one<N, 32, false>(p, stream);
}

}  // namespace svae
