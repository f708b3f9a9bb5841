// One translation unit per latent dimension for the VJP kernel (compiled with -DSVAE_N=<n>).
#include "lds_estep_bwd.cuh"
#include "lds_estep_bwd_packed.cuh"
#include "launch.h"
#include <cstdlib>

#ifndef SVAE_N
#error "compile with -DSVAE_N=<latent dim>"
#endif

#define SVAE_CAT_(a, b) a##b
#define SVAE_CAT(a, b) SVAE_CAT_(a, b)

#ifdef SVAE_STUB
// experiment builds (SVAE_BUILD_NS=...): latent dimensions left out of a variant library
namespace svae {
cudaError_t SVAE_CAT(lds_bwd_launch_n, SVAE_N)(const LdsBwdParams&, int, cudaStream_t) { return cudaErrorNotSupported; }
}
#else

#ifndef SVAE_LDS_BWD_PACKED_MAXN
#define SVAE_LDS_BWD_PACKED_MAXN 16
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

// the packed mapping (lds_estep_bwd_packed.cuh) covers the cotangents an SVI step produces
template <int N, int S>
static cudaError_t launch_bwd_packed(const LdsBwdParams& p, cudaStream_t stream) {
  constexpr int SPW = lds_packed_spw(N);
  const long long grid = ((long long)p.B + SPW - 1) / SPW;
  const size_t smem = sizeof(double) * (2 + (size_t)SPW * lds_bwd_packed_tile_doubles(N, S));
  auto kern = lds_estep_bwd_packed_kernel<N, S>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)grid, 32, smem, stream>>>(p);
  return cudaGetLastError();
}

static bool bwd_packed_eligible(const LdsBwdParams& p) {
  static const bool enabled = [] { const char* e = getenv("SVAE_LDS_PACKED"); return !(e && atoi(e) == 0); }();
  // largest latent dimension the packed VJP takes (it spills beyond n = 12: DESIGN.md 3.2); dev override by environment
  static const int max_n = [] { const char* e = getenv("SVAE_LDS_BWD_PACKED_MAXN"); return e ? atoi(e) : SVAE_LDS_BWD_PACKED_MAXN; }();
  if (SVAE_N > max_n) return false;
  return enabled && p.S <= 2 && (p.flags & (SVAE_LDS_BWD_SWEEP1_ONLY | SVAE_LDS_BWD_SWEEP2_ONLY)) == 0 && !p.g_E_init_xx &&
         !p.g_E_init_x && !p.g_E_pair_xx && !p.g_E_pair_xxn && !p.g_E_pair_xnxn && !p.g_Exx_t && !p.g_Exxn_t &&
         !p.g_node_J_dense && !p.g_msg_M && p.g_node_J && p.g_node_h;
}

cudaError_t SVAE_CAT(lds_bwd_launch_n, SVAE_N)(const LdsBwdParams& p, int lps, cudaStream_t stream) {
  constexpr int N = SVAE_N;
  if (bwd_packed_eligible(p)) {
    switch (p.S) {
      case 0: return launch_bwd_packed<N, 0>(p, stream);
      case 1: return launch_bwd_packed<N, 1>(p, stream);
      default: return launch_bwd_packed<N, 2>(p, stream);
    }
  }
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
#endif  // SVAE_STUB
