// Host build of svae_b200/csrc/lds_compose_core.h (TEST INFRASTRUCTURE): the same phase code the CUDA warp runs,
// with the lanes looped, so the arithmetic of svae_lds_compose_chunks is pinned on CPU without a GPU.
#include <vector>
#include <algorithm>
#include "lds_compose_core.h"

extern "C" int compose_chunks_host(int B, int T, int n, int chunk_len, const double* J11, const double* J12,
                                   const double* J22, const double* pair_logZ, long long stride_t, long long stride_b,
                                   const double* node_J,
                                   const double* node_h, const double* node_logZ, double* cJ11, double* cJ12,
                                   double* cJ22, double* clogZ, double* ce1, double* ce2) {
  const int C = (T - 1 + chunk_len - 1) / chunk_len;
  std::vector<double> mem(svae::compose_tile_doubles(n));
  int status = 0;
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const long long tile = (long long)b * C + c;
      const int s = c * chunk_len, L = std::min(chunk_len, T - 1 - s);
      const size_t pb = (size_t)b * stride_b;
      const int bad = svae::compose_chunk_n(n, mem.data(), s, L, J11 + pb, J12 + pb, J22 + pb,
                                          pair_logZ ? pair_logZ + pb / (n * n) : nullptr, stride_t,
                                          node_J + (size_t)b * T * n, node_h + (size_t)b * T * n,
                                          node_logZ ? node_logZ + (size_t)b * T : nullptr, cJ11 + tile * n * n,
                                          cJ12 + tile * n * n, cJ22 + tile * n * n, clogZ + tile, ce1 + tile * n,
                                          ce2 + tile * n);
      if (bad && !status) status = b + 1;
    }
  return status;
}
