// slds_glue.cu -- the two contractions that couple the discrete and the continuous factor of the SLDS local mean
// field, batched over sequences and time (bandwidth-bound elementwise kernels; they replace two groups of tall-skinny
// cuBLAS GEMMs whose inner dimension is K <= 32).
//
// What they replace (reference, one sequence per call, NumPy tensordot over Python lists):
//   svae/models/slds_svae.py:92-103   get_var_lds_local_natparam : pair parameters at time t = sum_k E[z_{t+1}=k] pair_k,
//                                      initial-state parameters    = sum_k E[z_0=k] init_k
//   svae/models/slds_svae.py:131-147  get_arhmm_local_nodeparams : HMM node potential (t,k) = <E pair stats_t, pair_k>
//                                      (+ <E init stats, init_k> at t = 0)
// Layout: one warp per (sequence, time) row; the K per-mode parameter blocks live in shared memory; a row's n^2-element
// matrices are read / written with consecutive lanes on consecutive elements (coalesced, conflict-free).
// `active` (optional, (B,) int32): sequences whose mean field has converged are skipped -- their outputs keep the
// values of the sweep that converged them (the reference stops iterating that sequence, slds_svae.py:171-172).
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "svae_b200.h"

extern std::atomic<uint64_t> g_svae_launches;

namespace svae {

constexpr int kSldsMaxK = 32;

// node (B,T,K).  Exx_t (B,T,n,n) = E[x_t x_t'], Exxn_t (B,T-1,n,n) = E[x_t x_{t+1}'], Ex0 (B,n) = E[x_0]
template <int KMAX>
__global__ void __launch_bounds__(256) slds_node_potentials_kernel(
    int B, int T, int n, int K, const double* __restrict__ Exx_t, const double* __restrict__ Exxn_t,
    const double* __restrict__ Ex0, const double* __restrict__ iJ, const double* __restrict__ ih,
    const double* __restrict__ ic, const double* __restrict__ J11, const double* __restrict__ J12,
    const double* __restrict__ J22, const double* __restrict__ lZ, const int* __restrict__ active,
    double* __restrict__ node) {
  extern __shared__ double sm[];
  const int nn = n * n;
  double* sJ = sm;                 // [k][e]  init precision
  double* s11 = sJ + K * nn;
  double* s12 = s11 + K * nn;
  double* s22 = s12 + K * nn;
  double* sh = s22 + K * nn;       // [k][i]
  for (int e = threadIdx.x; e < K * nn; e += blockDim.x) {
    sJ[e] = iJ[e]; s11[e] = J11[e]; s12[e] = J12[e]; s22[e] = J22[e];
  }
  for (int e = threadIdx.x; e < K * n; e += blockDim.x) sh[e] = ih[e];
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long rows = (long long)B * T;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const long long b = row / T;
    const int t = (int)(row - b * T);
    if (active && !active[b]) continue;
    double acc[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) acc[k] = 0.0;
    if (t == 0) {
      const double* a = Exx_t + (size_t)row * nn;
      for (int e = lane; e < nn; e += 32) {
        const double v = a[e];
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) acc[k] = fma(v, sJ[k * nn + e], acc[k]);
      }
      for (int e = lane; e < n; e += 32) {
        const double v = Ex0[(size_t)b * n + e];
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) acc[k] = fma(v, sh[k * n + e], acc[k]);
      }
    } else {
      const double* a = Exx_t + (size_t)(row - 1) * nn;                    // E[x_{t-1} x_{t-1}']
      const double* c = Exx_t + (size_t)row * nn;                          // E[x_t x_t']
      const double* x = Exxn_t + ((size_t)b * (T - 1) + (t - 1)) * nn;     // E[x_{t-1} x_t']
      for (int e = lane; e < nn; e += 32) {
        const double va = a[e], vx = x[e], vc = c[e];
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) acc[k] = fma(vc, s22[k * nn + e], fma(vx, s12[k * nn + e], fma(va, s11[k * nn + e], acc[k])));
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) node[(size_t)row * K + k] = v + (t == 0 ? ic[k] : lZ[k]);
      }
    }
  }
}

// init (B,n,n), (B,n), (B) and pair (B,T-1,n,n) x3, (B,T-1) from E[z] (B,T,K): row r = b*T + t, t = 0 mixes the
// initial-state parameters with E[z_0], t >= 1 the transition into t with E[z_t]
template <int KMAX>
__global__ void __launch_bounds__(256) slds_mix_kernel(
    int B, int T, int n, int K, const double* __restrict__ Ez, const double* __restrict__ iJ,
    const double* __restrict__ ih, const double* __restrict__ ic, const double* __restrict__ J11,
    const double* __restrict__ J12, const double* __restrict__ J22, const double* __restrict__ lZ,
    const int* __restrict__ active, double* __restrict__ oJ, double* __restrict__ oh, double* __restrict__ oc,
    double* __restrict__ o11, double* __restrict__ o12, double* __restrict__ o22, double* __restrict__ olZ) {
  extern __shared__ double sm[];
  const int nn = n * n;
  double* sJ = sm;
  double* s11 = sJ + K * nn;
  double* s12 = s11 + K * nn;
  double* s22 = s12 + K * nn;
  double* sh = s22 + K * nn;
  double* sc = sh + K * n;         // [k]: init constant, then pair logZ
  for (int e = threadIdx.x; e < K * nn; e += blockDim.x) {
    sJ[e] = iJ[e]; s11[e] = J11[e]; s12[e] = J12[e]; s22[e] = J22[e];
  }
  for (int e = threadIdx.x; e < K * n; e += blockDim.x) sh[e] = ih[e];
  for (int e = threadIdx.x; e < K; e += blockDim.x) { sc[e] = ic[e]; sc[K + e] = lZ[e]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long rows = (long long)B * T;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const long long b = row / T;
    const int t = (int)(row - b * T);
    if (active && !active[b]) continue;
    double w[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) w[k] = (k < K) ? Ez[(size_t)row * K + k] : 0.0;
    if (t == 0) {
      for (int e = lane; e < nn; e += 32) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) v = fma(w[k], sJ[k * nn + e], v);
        oJ[(size_t)b * nn + e] = v;
      }
      for (int e = lane; e < n; e += 32) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) v = fma(w[k], sh[k * n + e], v);
        oh[(size_t)b * n + e] = v;
      }
      if (lane == 0) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) v = fma(w[k], sc[k], v);
        oc[b] = v;
      }
    } else {
      const size_t o = ((size_t)b * (T - 1) + (t - 1)) * nn;
      for (int e = lane; e < nn; e += 32) {
        double v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < K) {
            v1 = fma(w[k], s11[k * nn + e], v1);
            v2 = fma(w[k], s12[k * nn + e], v2);
            v3 = fma(w[k], s22[k * nn + e], v3);
          }
        }
        o11[o + e] = v1; o12[o + e] = v2; o22[o + e] = v3;
      }
      if (lane == 0) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) v = fma(w[k], sc[K + k], v);
        olZ[(size_t)b * (T - 1) + (t - 1)] = v;
      }
    }
  }
}

size_t glue_smem(int n, int K) { return sizeof(double) * ((size_t)4 * K * n * n + (size_t)K * n + 2 * K); }

size_t glue_smem(int n, int K);
template <typename Kern, typename... Args>
static cudaError_t glue_launch(Kern kern, int B, int T, int n, int K, cudaStream_t stream, Args... args) {
  const size_t smem = glue_smem(n, K);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long rows = (long long)B * T;
  long long grid = (rows + 7) / 8;
  const long long cap = (long long)sms * 8;       // a multiple of the SM count; rows are strided over the grid
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, 256, smem, stream>>>(B, T, n, K, args...);
  return cudaGetLastError();
}

}  // namespace svae

extern "C" int svae_slds_node_potentials(int B, int T, int n, int K, const double* Exx_t, const double* Exxn_t,
                                         const double* Ex0, const double* init_J, const double* init_h,
                                         const double* init_c, const double* J11, const double* J12, const double* J22,
                                         const double* pair_logZ, const int* active, double* node, void* stream) {
  if (B <= 0 || T <= 0 || n <= 0 || K <= 0) return SVAE_ERR_BAD_ARG;
  if (!Exx_t || !Ex0 || !init_J || !init_h || !init_c || !node) return SVAE_ERR_BAD_ARG;
  if (T > 1 && (!Exxn_t || !J11 || !J12 || !J22 || !pair_logZ)) return SVAE_ERR_BAD_ARG;
  if (K > svae::kSldsMaxK || n > SVAE_LDS_MAX_N || svae::glue_smem(n, K) > 200 * 1024) return SVAE_ERR_UNSUPPORTED;
  if (T == 1) { J11 = J12 = J22 = init_J; pair_logZ = init_c; Exxn_t = Exx_t; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (K <= 8) e = svae::glue_launch(svae::slds_node_potentials_kernel<8>, B, T, n, K, st, Exx_t, Exxn_t, Ex0, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, node);
  else if (K <= 16) e = svae::glue_launch(svae::slds_node_potentials_kernel<16>, B, T, n, K, st, Exx_t, Exxn_t, Ex0, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, node);
  else e = svae::glue_launch(svae::slds_node_potentials_kernel<32>, B, T, n, K, st, Exx_t, Exxn_t, Ex0, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, node);
  g_svae_launches.fetch_add(1);
  return e == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}

extern "C" int svae_slds_mix_params(int B, int T, int n, int K, const double* Ez, const double* init_J,
                                    const double* init_h, const double* init_c, const double* J11, const double* J12,
                                    const double* J22, const double* pair_logZ, const int* active, double* out_init_J,
                                    double* out_init_h, double* out_init_c, double* out_J11, double* out_J12,
                                    double* out_J22, double* out_pair_logZ, void* stream) {
  if (B <= 0 || T <= 0 || n <= 0 || K <= 0) return SVAE_ERR_BAD_ARG;
  if (!Ez || !init_J || !init_h || !init_c || !out_init_J || !out_init_h || !out_init_c) return SVAE_ERR_BAD_ARG;
  if (T > 1 && (!J11 || !J12 || !J22 || !pair_logZ || !out_J11 || !out_J12 || !out_J22 || !out_pair_logZ)) return SVAE_ERR_BAD_ARG;
  if (K > svae::kSldsMaxK || n > SVAE_LDS_MAX_N || svae::glue_smem(n, K) > 200 * 1024) return SVAE_ERR_UNSUPPORTED;
  if (T == 1) { J11 = J12 = J22 = init_J; pair_logZ = init_c; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (K <= 8) e = svae::glue_launch(svae::slds_mix_kernel<8>, B, T, n, K, st, Ez, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, out_init_J, out_init_h, out_init_c, out_J11, out_J12, out_J22, out_pair_logZ);
  else if (K <= 16) e = svae::glue_launch(svae::slds_mix_kernel<16>, B, T, n, K, st, Ez, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, out_init_J, out_init_h, out_init_c, out_J11, out_J12, out_J22, out_pair_logZ);
  else e = svae::glue_launch(svae::slds_mix_kernel<32>, B, T, n, K, st, Ez, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, out_init_J, out_init_h, out_init_c, out_J11, out_J12, out_J22, out_pair_logZ);
  g_svae_launches.fetch_add(1);
  return e == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}
