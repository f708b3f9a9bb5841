// reduce_stats.cu -- minibatch reduction of the LDS global expected statistics (deterministic, two stages).
//
// The SVI step needs sum_b of (E_init, E_pair) and of lognorm (svae/svae.py:33-34: flat(stats) of ONE
// minibatch); with the batch axis this build adds, that is a (B, 4n^2+n+1) -> (4n^2+n+1) column sum.  The
// result is the buffer that the single NCCL all-reduce per step carries across GPUs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "svae_b200.h"

extern std::atomic<uint64_t> g_svae_launches;

namespace {

constexpr int kStage1Ctas = 148;

struct ReduceParams {
  int B, n;
  const double *E_init_xx, *E_init_x, *E_pair_xx, *E_pair_xxn, *E_pair_xnxn, *lognorm;
  double* partial;   // (kStage1Ctas, E)
  double* out;       // (E)
};

__device__ __forceinline__ double load_elem(const ReduceParams& p, long long b, int e) {
  const int nn = p.n * p.n;
  if (e < nn) return p.E_init_xx[b * nn + e];
  e -= nn;
  if (e < p.n) return p.E_init_x[b * p.n + e];
  e -= p.n;
  if (e < nn) return p.E_pair_xx[b * nn + e];
  e -= nn;
  if (e < nn) return p.E_pair_xxn[b * nn + e];
  e -= nn;
  if (e < nn) return p.E_pair_xnxn[b * nn + e];
  return p.lognorm[b];
}

__global__ void reduce_stage1(const ReduceParams p) {
  const int E = 4 * p.n * p.n + p.n + 1;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    double acc = 0.0;
    for (long long b = blockIdx.x; b < p.B; b += gridDim.x) acc += load_elem(p, b, e);
    p.partial[(size_t)blockIdx.x * E + e] = acc;
  }
}

// one warp per output element: lanes stride over the partials, then a fixed-order shuffle tree
__global__ void reduce_stage2(const ReduceParams p, int parts) {
  const int E = 4 * p.n * p.n + p.n + 1;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  double acc = 0.0;
  for (int g = lane; g < parts; g += 32) acc += p.partial[(size_t)g * E + e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) p.out[e] = acc;
}

}  // namespace

extern "C" size_t svae_lds_reduce_stats_workspace_bytes(int n) {
  if (n <= 0) return 0;
  return sizeof(double) * (size_t)kStage1Ctas * (size_t)(4 * n * n + n + 1);
}

extern "C" int svae_lds_reduce_stats(int B, int n, const double* E_init_xx, const double* E_init_x,
                                     const double* E_pair_xx, const double* E_pair_xxn, const double* E_pair_xnxn,
                                     const double* lognorm, double* out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (B <= 0 || n <= 0) return SVAE_ERR_BAD_ARG;
  if (!E_init_xx || !E_init_x || !E_pair_xx || !E_pair_xxn || !E_pair_xnxn || !lognorm || !out || !workspace)
    return SVAE_ERR_BAD_ARG;
  if (workspace_bytes < svae_lds_reduce_stats_workspace_bytes(n)) return SVAE_ERR_WORKSPACE;
  ReduceParams p;
  p.B = B; p.n = n;
  p.E_init_xx = E_init_xx; p.E_init_x = E_init_x; p.E_pair_xx = E_pair_xx; p.E_pair_xxn = E_pair_xxn;
  p.E_pair_xnxn = E_pair_xnxn; p.lognorm = lognorm;
  p.partial = static_cast<double*>(workspace); p.out = out;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int parts = B < kStage1Ctas ? B : kStage1Ctas;
  reduce_stage1<<<parts, 256, 0, s>>>(p);
  const int E = 4 * n * n + n + 1;
  reduce_stage2<<<(E + 7) / 8, 256, 0, s>>>(p, parts);
  g_svae_launches.fetch_add(2);
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}
