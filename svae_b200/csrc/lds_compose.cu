// lds_compose.cu -- svae_lds_compose_chunks: one warp per (sequence, chunk) integrates a chunk of the LDS chain into a
// single pair potential (lds_compose_core.h).  Stage 1 of the parallel-in-time E-step for the small-batch / long-T
// regime; stages 2 and 3 are ordinary launches of the fused E-step kernel.
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include "svae_b200.h"
#include "lds_compose_core.h"

extern std::atomic<uint64_t> g_svae_launches;

namespace {
constexpr int kWarps = 4;

struct ComposeParams {
  int B, T, n, Lc, C;
  long long stride_t, stride_b;
  const double *J11, *J12, *J22, *pair_logZ, *node_J, *node_h, *node_logZ;
  double *cJ11, *cJ12, *cJ22, *clogZ, *ce1, *ce2;
  int* status;
};

template <int N>
__global__ void __launch_bounds__(32 * kWarps) lds_compose_kernel(const ComposeParams p) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5;
  const long long tile = (long long)blockIdx.x * kWarps + warp;
  if (tile >= (long long)p.B * p.C) return;           // warp-uniform
  const int b = (int)(tile / p.C), c = (int)(tile % p.C);
  const int s = c * p.Lc;
  const int L = min(p.Lc, p.T - 1 - s);
  constexpr int n = N, nn = N * N;
  const size_t bt = (size_t)b * p.T;
  const size_t pb = (size_t)b * p.stride_b;
  const int bad = svae::compose_chunk<N>(
      smem + (size_t)warp * svae::compose_tile_doubles(n), s, L, p.J11 + pb, p.J12 + pb, p.J22 + pb,
      p.pair_logZ ? p.pair_logZ + pb / nn : nullptr, p.stride_t, p.node_J + bt * n, p.node_h + bt * n,
      p.node_logZ ? p.node_logZ + bt : nullptr, p.cJ11 + tile * nn, p.cJ12 + tile * nn, p.cJ22 + tile * nn, p.clogZ + tile,
      p.ce1 + tile * n, p.ce2 + tile * n);
  if (bad && p.status && (threadIdx.x & 31) == 0) atomicCAS(p.status, 0, b + 1);
}

template <int N>
cudaError_t launch_compose(const ComposeParams& p, cudaStream_t stream) {
  const size_t smem = sizeof(double) * (size_t)kWarps * svae::compose_tile_doubles(N);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lds_compose_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const long long tiles = (long long)p.B * p.C;
  lds_compose_kernel<N><<<(unsigned)((tiles + kWarps - 1) / kWarps), 32 * kWarps, smem, stream>>>(p);
  return cudaGetLastError();
}
typedef cudaError_t (*compose_fn)(const ComposeParams&, cudaStream_t);
const compose_fn kCompose[SVAE_LDS_MAX_N + 1] = {
    nullptr, launch_compose<1>, launch_compose<2>, launch_compose<3>, launch_compose<4>, launch_compose<5>,
    launch_compose<6>, launch_compose<7>, launch_compose<8>, launch_compose<9>, launch_compose<10>, launch_compose<11>,
    launch_compose<12>, launch_compose<13>, launch_compose<14>, launch_compose<15>, launch_compose<16>};
}  // namespace

extern "C" int svae_lds_compose_num_chunks(int T, int chunk_len) {
  if (T < 2 || chunk_len < 1) return 0;
  return (T - 1 + chunk_len - 1) / chunk_len;
}

extern "C" int svae_lds_compose_chunks(int B, int T, int n, int chunk_len, const double* J11, const double* J12,
                                       const double* J22, const double* pair_logZ, long long pair_stride_t,
                                       long long pair_stride_b, const double* node_J,
                                       const double* node_h, const double* node_logZ, double* chunk_J11,
                                       double* chunk_J12, double* chunk_J22, double* chunk_logZ, double* chunk_e1,
                                       double* chunk_e2, int* status, void* stream) {
  if (B <= 0 || T < 2 || n <= 0 || chunk_len < 1 || pair_stride_t < 0 || pair_stride_b < 0) return SVAE_ERR_BAD_ARG;
  if (pair_stride_t % ((long long)n * n) != 0 || pair_stride_b % ((long long)n * n) != 0) return SVAE_ERR_BAD_ARG;
  if (!J11 || !J12 || !J22 || !node_J || !node_h || !chunk_J11 || !chunk_J12 || !chunk_J22 || !chunk_logZ || !chunk_e1 ||
      !chunk_e2)
    return SVAE_ERR_BAD_ARG;
  if (n > SVAE_LDS_MAX_N) return SVAE_ERR_UNSUPPORTED;
  ComposeParams p;
  p.B = B; p.T = T; p.n = n; p.Lc = chunk_len; p.C = svae_lds_compose_num_chunks(T, chunk_len);
  p.J11 = J11; p.J12 = J12; p.J22 = J22; p.pair_logZ = pair_logZ;
  p.stride_t = pair_stride_t; p.stride_b = pair_stride_b;
  p.node_J = node_J; p.node_h = node_h; p.node_logZ = node_logZ;
  p.cJ11 = chunk_J11; p.cJ12 = chunk_J12; p.cJ22 = chunk_J22; p.clogZ = chunk_logZ; p.ce1 = chunk_e1; p.ce2 = chunk_e2;
  p.status = status;
  const cudaError_t e = kCompose[n](p, static_cast<cudaStream_t>(stream));
  g_svae_launches.fetch_add(1);
  return e == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}
