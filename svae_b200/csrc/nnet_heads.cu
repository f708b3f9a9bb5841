// nnet_heads.cu -- the recognition network's output head fused with node-potential production.
//
// What it replaces (reference, NumPy under autograd): svae/nnet.py:43-47 gaussian_info (J = -1/2 log1pexp(J_in), h) and the
// 'info' branch of the Gaussian resnet svae/nnet.py:90-104 (J += -1/2 log1pexp(b2), h += inputs W + b1): five elementwise
// passes over the (rows, 2n) activation become one kernel that writes the node potentials (J_node, h_node) of the E-step
// directly, and one kernel for its VJP.  Memory-bound elementwise work: one thread per output element, coalesced.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <atomic>
#include "svae_b200.h"

extern std::atomic<uint64_t> g_svae_launches;

namespace svae {

__device__ __forceinline__ double softplus(double x) { return fmax(x, 0.0) + log1p(exp(-fabs(x))); }
__device__ __forceinline__ double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

// act (R, 2n) = [J_in | h_in];  res (R, n) = inputs W + b1 (may be NULL);  b2 (n) (may be NULL)
__global__ void __launch_bounds__(256) info_head_fwd_kernel(long long R, int n, const double* __restrict__ act,
                                                            const double* __restrict__ res, const double* __restrict__ b2,
                                                            double* __restrict__ J, double* __restrict__ h) {
  const long long total = R * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / n;
    const int i = (int)(e - r * n);
    const double a = act[r * 2 * n + i];
    J[e] = -0.5 * (softplus(a) + (b2 ? softplus(b2[i]) : 0.0));
    h[e] = act[r * 2 * n + n + i] + (res ? res[e] : 0.0);
  }
}

// g_act (R, 2n) = [-1/2 sigmoid(J_in) gJ | gh]
__global__ void __launch_bounds__(256) info_head_bwd_kernel(long long R, int n, const double* __restrict__ act,
                                                            const double* __restrict__ gJ, const double* __restrict__ gh,
                                                            double* __restrict__ g_act) {
  const long long total = R * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / n;
    const int i = (int)(e - r * n);
    g_act[r * 2 * n + i] = gJ ? -0.5 * sigmoid(act[r * 2 * n + i]) * gJ[e] : 0.0;
    g_act[r * 2 * n + n + i] = gh ? gh[e] : 0.0;
  }
}

static unsigned grid_for(long long total) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long g = (total + 255) / 256;
  const long long cap = (long long)sms * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace svae

extern "C" int svae_gaussian_info_head_fwd(long long rows, int n, const double* act, const double* res, const double* b2,
                                           double* node_J, double* node_h, void* stream) {
  if (rows <= 0 || n <= 0 || !act || !node_J || !node_h) return SVAE_ERR_BAD_ARG;
  svae::info_head_fwd_kernel<<<svae::grid_for(rows * n), 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, n, act, res, b2, node_J, node_h);
  g_svae_launches.fetch_add(1);
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}

extern "C" int svae_gaussian_info_head_bwd(long long rows, int n, const double* act, const double* g_node_J,
                                           const double* g_node_h, double* g_act, void* stream) {
  if (rows <= 0 || n <= 0 || !act || !g_act) return SVAE_ERR_BAD_ARG;
  svae::info_head_bwd_kernel<<<svae::grid_for(rows * n), 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, n, act, g_node_J, g_node_h, g_act);
  g_svae_launches.fetch_add(1);
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}
