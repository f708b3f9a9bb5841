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
#include <cstdlib>
#include <initializer_list>
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


// ------------------------------------------------------------------------------------------------------------------
// K <= 8: both contractions are (B T) x {3 n^2} x 8 products in fp64 -- the shape of the fp64 tensor-core instruction
// mma.m8n8k4 (one warp: 8 rows x 8 modes, 4 elements of the inner dimension per issue).  A warp owns 8 consecutive
// (sequence, time) rows; the K mode blocks sit in shared memory in the fragment order, so a step is one global (or one
// shared) load + one mma per lane instead of 24 shared loads + 24 fma of the row-per-warp kernels below, which stay as
// the general path (K <= 32).  Fragment layout (PTX ISA, mma.m8n8k4 .f64): g = lane >> 2, q = lane & 3;
// A[g][q], B[q][g], C[g][2q], C[g][2q+1].
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__host__ __device__ inline int mma_nn4(int n) { return (n * n + 3) & ~3; }     // inner dimension padded to the k-step
__host__ __device__ inline int mma_n4(int n) { return (n + 3) & ~3; }
// mix kernel: row stride of the [k][e] blocks, == 8 (mod 16) doubles so that the 4 k-rows of a B fragment tile the banks
__host__ __device__ inline int mma_ld8(int cols) { const int c = (cols + 7) & ~7; return (c & 15) == 8 ? c : c + 8; }

size_t glue_mma_node_smem(int n) { return sizeof(double) * 8 * ((size_t)4 * mma_nn4(n) + mma_n4(n)); }
size_t glue_mma_mix_smem(int n) { return sizeof(double) * 8 * ((size_t)4 * mma_ld8(n * n) + mma_ld8(n + 2)); }

// node (B,T,8-or-fewer): see slds_node_potentials_kernel for the contraction
__global__ void __launch_bounds__(256) slds_node_potentials_mma_kernel(
    int B, int T, int n, int K, const double* __restrict__ Exx_t, const double* __restrict__ Exxn_t,
    const double* __restrict__ Ex0, const double* __restrict__ iJ, const double* __restrict__ ih,
    const double* __restrict__ ic, const double* __restrict__ J11, const double* __restrict__ J12,
    const double* __restrict__ J22, const double* __restrict__ lZ, const int* __restrict__ active,
    double* __restrict__ node) {
  extern __shared__ double sm[];
  const int nn = n * n, nn4 = mma_nn4(n), n4 = mma_n4(n);
  double* s11 = sm;                   // [e][k], k fastest (8 wide, zero beyond K and beyond n^2)
  double* s12 = s11 + 8 * nn4;
  double* s22 = s12 + 8 * nn4;
  double* sJ = s22 + 8 * nn4;
  double* sh = sJ + 8 * nn4;          // [i][k]
  for (int x = threadIdx.x; x < 8 * nn4; x += blockDim.x) {
    const int e = x >> 3, k = x & 7;
    const bool in = e < nn && k < K;
    s11[x] = in ? J11[k * nn + e] : 0.0;
    s12[x] = in ? J12[k * nn + e] : 0.0;
    s22[x] = in ? J22[k * nn + e] : 0.0;
    sJ[x] = in ? iJ[k * nn + e] : 0.0;
  }
  for (int x = threadIdx.x; x < 8 * n4; x += blockDim.x) {
    const int e = x >> 3, k = x & 7;
    sh[x] = (e < n && k < K) ? ih[k * n + e] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, wpb = blockDim.x >> 5;
  const long long rows = (long long)B * T, groups = (rows + 7) >> 3;
  for (long long grp = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); grp < groups; grp += (long long)gridDim.x * wpb) {
    const long long row = grp * 8 + g;
    const bool inr = row < rows;
    const long long rc = inr ? row : rows - 1;
    const long long b = rc / T;
    const int t = (int)(rc - b * T);
    const bool act = inr && (!active || active[b] != 0);
    const bool trans = act && t > 0, first = act && t == 0;
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0, f0 = 0.0, f1 = 0.0;       // three independent accumulation chains
    const double* pa = Exx_t + (size_t)(rc - (t > 0)) * nn + q;            // E[x_{t-1} x_{t-1}']
    const double* pc = Exx_t + (size_t)rc * nn + q;                        // E[x_t x_t']
    const double* px = Exxn_t + ((size_t)b * (T - 1) + (t > 0 ? t - 1 : 0)) * nn + q;   // E[x_{t-1} x_t']
    const double* b11 = s11 + q * 8 + g;
    const double* b12 = s12 + q * 8 + g;
    const double* b22 = s22 + q * 8 + g;
#pragma unroll 5
    for (int e0 = 0; e0 < nn4; e0 += 4) {
      const bool ld = trans && e0 + q < nn;
      const double va = ld ? pa[e0] : 0.0, vx = ld ? px[e0] : 0.0, vc = ld ? pc[e0] : 0.0;
      dmma884(c0, c1, va, b11[e0 * 8]);
      dmma884(d0, d1, vx, b12[e0 * 8]);
      dmma884(f0, f1, vc, b22[e0 * 8]);
    }
    if (__any_sync(0xffffffffu, first)) {                                   // a row group holding the start of a sequence
      const double* bJ = sJ + q * 8 + g;
      const double* bh = sh + q * 8 + g;
      for (int e0 = 0; e0 < nn4; e0 += 4) dmma884(c0, c1, (first && e0 + q < nn) ? pc[e0] : 0.0, bJ[e0 * 8]);
      const double* p0 = Ex0 + (size_t)b * n + q;
      for (int e0 = 0; e0 < n4; e0 += 4) dmma884(d0, d1, (first && e0 + q < n) ? p0[e0] : 0.0, bh[e0 * 8]);
    }
    if (act) {
      const int k0 = 2 * q;
      const double* cst = (t == 0) ? ic : lZ;
      double* o = node + (size_t)row * K + k0;
      const double v0 = c0 + d0 + f0, v1 = c1 + d1 + f1;
      if (K == 8) {
        *reinterpret_cast<double2*>(o) = make_double2(v0 + cst[k0], v1 + cst[k0 + 1]);
      } else {
        if (k0 < K) o[0] = v0 + cst[k0];
        if (k0 + 1 < K) o[1] = v1 + cst[k0 + 1];
      }
    }
  }
}

// init / pair parameters mixed with E[z] (B,T,K <= 8): see slds_mix_kernel
__global__ void __launch_bounds__(256) slds_mix_mma_kernel(
    int B, int T, int n, int K, const double* __restrict__ Ez, const double* __restrict__ iJ,
    const double* __restrict__ ih, const double* __restrict__ ic, const double* __restrict__ J11,
    const double* __restrict__ J12, const double* __restrict__ J22, const double* __restrict__ lZ,
    const int* __restrict__ active, double* __restrict__ oJ, double* __restrict__ oh, double* __restrict__ oc,
    double* __restrict__ o11, double* __restrict__ o12, double* __restrict__ o22, double* __restrict__ olZ) {
  extern __shared__ double sm[];
  const int nn = n * n, ld = mma_ld8(nn), ldv = mma_ld8(n + 2);
  double* s11 = sm;                  // [k][e], e fastest, row stride ld (zero beyond K and beyond n^2)
  double* s12 = s11 + 8 * ld;
  double* s22 = s12 + 8 * ld;
  double* sJ = s22 + 8 * ld;
  double* sv = sJ + 8 * ld;          // [k][ ih (n) | ic | lZ ]
  for (int x = threadIdx.x; x < 8 * ld; x += blockDim.x) {
    const int k = x / ld, e = x - k * ld;
    const bool in = e < nn && k < K;
    s11[x] = in ? J11[k * nn + e] : 0.0;
    s12[x] = in ? J12[k * nn + e] : 0.0;
    s22[x] = in ? J22[k * nn + e] : 0.0;
    sJ[x] = in ? iJ[k * nn + e] : 0.0;
  }
  for (int x = threadIdx.x; x < 8 * ldv; x += blockDim.x) {
    const int k = x / ldv, e = x - k * ldv;
    sv[x] = (k >= K) ? 0.0 : (e < n ? ih[k * n + e] : (e == n ? ic[k] : (e == n + 1 ? lZ[k] : 0.0)));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, wpb = blockDim.x >> 5;
  const long long rows = (long long)B * T, groups = (rows + 7) >> 3;
  const bool vec2 = (nn & 1) == 0;                                          // 16-byte stores stay aligned
  for (long long grp = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); grp < groups; grp += (long long)gridDim.x * wpb) {
    const long long rowA = grp * 8 + g;                                     // A fragment: this lane's row of E[z]
    const long long rcA = rowA < rows ? rowA : rows - 1;
    const double w_lo = (rowA < rows && q < K) ? Ez[(size_t)rcA * K + q] : 0.0;
    const double w_hi = (rowA < rows && 4 + q < K) ? Ez[(size_t)rcA * K + 4 + q] : 0.0;
    // C fragment: row g again, columns e0 + 2q, e0 + 2q + 1
    const long long b = rcA / T;
    const int t = (int)(rcA - b * T);
    const bool act = rowA < rows && (!active || active[b] != 0);
    const bool trans = act && t > 0, first = act && t == 0;
    const size_t o = ((size_t)b * (T - 1) + (t > 0 ? t - 1 : 0)) * nn;
    auto tile = [&](const double* sB, int stride, int e0, double& v0, double& v1) {
      v0 = 0.0; v1 = 0.0;
      dmma884(v0, v1, w_lo, sB[q * stride + e0 + g]);
      dmma884(v0, v1, w_hi, sB[(4 + q) * stride + e0 + g]);
    };
    auto put = [&](double* dst, int e, int lim, double v0, double v1) {       // dst already offset to the row
      if (vec2 && e + 1 < lim) *reinterpret_cast<double2*>(dst + e) = make_double2(v0, v1);
      else { if (e < lim) dst[e] = v0; if (e + 1 < lim) dst[e + 1] = v1; }
    };
    if (__any_sync(0xffffffffu, trans)) {
#pragma unroll 2
      for (int e0 = 0; e0 < nn; e0 += 8) {
        double a0, a1, x0, x1, c0, c1;
        tile(s11, ld, e0, a0, a1);
        tile(s12, ld, e0, x0, x1);
        tile(s22, ld, e0, c0, c1);
        if (trans) {
          const int e = e0 + 2 * q;
          put(o11 + o, e, nn, a0, a1);
          put(o12 + o, e, nn, x0, x1);
          put(o22 + o, e, nn, c0, c1);
        }
      }
    }
    if (__any_sync(0xffffffffu, first)) {
      for (int e0 = 0; e0 < nn; e0 += 8) {
        double v0, v1;
        tile(sJ, ld, e0, v0, v1);
        if (first) put(oJ + (size_t)b * nn, e0 + 2 * q, nn, v0, v1);
      }
    }
    for (int e0 = 0; e0 < n + 2; e0 += 8) {                                   // h (t = 0), init constant (t = 0), pair logZ (t > 0)
      double v[2];
      tile(sv, ldv, e0, v[0], v[1]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int e = e0 + 2 * q + j;
        if (first && e < n) oh[(size_t)b * n + e] = v[j];
        if (first && e == n) oc[b] = v[j];
        if (trans && e == n + 1) olZ[(size_t)b * (T - 1) + (t - 1)] = v[j];
      }
    }
  }
}

template <typename Kern, typename... Args>
static cudaError_t glue_mma_launch(Kern kern, size_t smem, int B, int T, int n, int K, cudaStream_t stream, Args... args) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long groups = ((long long)B * T + 7) / 8;
  long long grid = (groups + 7) / 8;                 // 8 warps per CTA, one 8-row group per warp and pass
  const long long cap = (long long)sms * 8;
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, 256, smem, stream>>>(B, T, n, K, args...);
  return cudaGetLastError();
}

static bool aligned16(std::initializer_list<const void*> ps) {
  for (const void* q : ps) if (reinterpret_cast<uintptr_t>(q) & 15) return false;
  return true;
}

static bool glue_use_mma(int n, int K) {
  static const bool enabled = [] { const char* e = getenv("SVAE_SLDS_MMA"); return !(e && atoi(e) == 0); }();
  return enabled && K <= 8 && glue_mma_node_smem(n) <= 200 * 1024 && glue_mma_mix_smem(n) <= 200 * 1024;
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
  if (svae::glue_use_mma(n, K) && svae::aligned16({node})) e = svae::glue_mma_launch(svae::slds_node_potentials_mma_kernel, svae::glue_mma_node_smem(n), B, T, n, K, st, Exx_t, Exxn_t, Ex0, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, node);
  else if (K <= 8) e = svae::glue_launch(svae::slds_node_potentials_kernel<8>, B, T, n, K, st, Exx_t, Exxn_t, Ex0, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, node);
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
  if (svae::glue_use_mma(n, K) && svae::aligned16({out_init_J, out_J11, out_J12, out_J22})) e = svae::glue_mma_launch(svae::slds_mix_mma_kernel, svae::glue_mma_mix_smem(n), B, T, n, K, st, Ez, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, out_init_J, out_init_h, out_init_c, out_J11, out_J12, out_J22, out_pair_logZ);
  else if (K <= 8) e = svae::glue_launch(svae::slds_mix_kernel<8>, B, T, n, K, st, Ez, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, out_init_J, out_init_h, out_init_c, out_J11, out_J12, out_J22, out_pair_logZ);
  else if (K <= 16) e = svae::glue_launch(svae::slds_mix_kernel<16>, B, T, n, K, st, Ez, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, out_init_J, out_init_h, out_init_c, out_J11, out_J12, out_J22, out_pair_logZ);
  else e = svae::glue_launch(svae::slds_mix_kernel<32>, B, T, n, K, st, Ez, init_J, init_h, init_c, J11, J12, J22, pair_logZ, active, out_init_J, out_init_h, out_init_c, out_J11, out_J12, out_J22, out_pair_logZ);
  g_svae_launches.fetch_add(1);
  return e == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}
