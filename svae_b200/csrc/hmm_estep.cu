// hmm_estep.cu -- batched HMM local E-step (forward-backward), log-normaliser and Viterbi for sm_100a.
//
// What it replaces (reference, one sequence per call):
//   svae/hmm/cython_hmm_inference.pyx:93-121   hmm_logZ       (log-space forward, T K^2 exp)
//   svae/hmm/cython_hmm_inference.pyx:126-166  hmm_logZ_grad  (reverse sweep; with g = 1 the E-step)
//   svae/hmm/cython_hmm_inference.pyx:59-88    hmm_logZ_normalized (the scaled-forward formulation)
//   svae/hmm/hmm_inference.py:21-41,54-63      hmm_estep / hmm_viterbi (pyhsmm log messages + Viterbi)
//
// B200 design: one sub-warp tile of LPS = 2^k >= K lanes per sequence, lane j = state j (32/LPS sequences
// per warp).  SCALED forward-backward (T K exps instead of T K^2, hmm_logZ_normalized's formulation):
//   forward : alpha_t  ~ (alpha_{t-1}' P) .* exp(node_t - max_t),  P = exp(log A - max log A) held as one
//             COLUMN per lane in registers, alpha broadcast through shared memory; logZ accumulates log c_t;
//             alpha_t is parked in the E[z_t] output buffer (no extra workspace);
//   backward: beta~_t ~ P (exp(node_{t+1} - max) .* beta~_{t+1}),  P held as one ROW per lane;
//             E[z_t] = alpha_t beta~_t / s_t,  s_t = <alpha_t, beta~_t>;
//             E[#i->j] = P[i][j] * sum_t alpha_t[i] w_{t+1}[j] / s_t  accumulated as one column per lane.
// Node potentials stream once forward and once backward with a one-step register prefetch.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <atomic>
#include "svae_b200.h"

namespace svae {

struct HmmParams {
  int B, T, K, flags;
  const double *log_pi, *log_A, *node;   // (K), (K,K) row = from, (B,T,K)
  double *logZ, *Ez0, *Etrans, *Ez;      // (B), (B,K), (B,K,K), (B,T,K)
  const int* active;                     // optional (B,) mask: sequences with active[b] == 0 are left untouched
};

template <int LPS>
__device__ __forceinline__ double tile_sum(double v) {
#pragma unroll
  for (int o = LPS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LPS);
  return v;
}
template <int LPS>
__device__ __forceinline__ float tile_maxf(float v) {
#pragma unroll
  for (int o = LPS / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, LPS));
  return v;
}

__device__ __forceinline__ double hmm_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// flags: bit0 = forward only (logZ), else full E-step.
//
// Latency structure: the per-step chain is  message -> shared memory -> K-term dot product -> next message.
// Every normaliser is taken OFF that chain: the sum of the previous message is re-derived by each lane from the
// very values it broadcasts for the dot product (no shuffle tree), is applied one step late (any positive
// rescaling is legal, the logs are accounted exactly), and exp(node - max) of the NEXT step is computed while
// the current dot product is in flight.  In the backward sweep E[z_t], the transition-count outer product and
// the normaliser 1/s_t of step t are finished during step t-1.
// -DSVAE_HMM_OSMEM=Q (K in 17..32 only): the transition-count accumulator (one 32-double column per lane, 64 registers)
// lives in shared memory and is updated every Q steps from a ring of the last Q alpha vectors, so the kernel fits 128
// registers and a K = 32 batch of 2048 sequences is resident in ONE wave (16 warps per SM) instead of two.
#ifndef SVAE_HMM_OSMEM
#define SVAE_HMM_OSMEM 0
#endif
#ifndef SVAE_HMM_MINB32
#define SVAE_HMM_MINB32 (SVAE_HMM_OSMEM ? 16 : 8)
#endif
template <int LPS>
__global__ void __launch_bounds__(32, (LPS == 32) ? SVAE_HMM_MINB32 : 16) hmm_estep_kernel(const HmmParams p) {
  constexpr int SPW = 32 / LPS;
  constexpr int LP = (LPS + 1) & ~1;
  constexpr int LQ = LP < 2 ? 2 : LP;
  __shared__ __align__(16) double sbuf[SPW][8][LQ];
  constexpr int OQ = (LPS == 32) ? SVAE_HMM_OSMEM : 0;       // pending outer products per flush (0: registers)
  __shared__ __align__(16) double Osm[OQ ? 32 * 32 : 2];      // Osm[i*32 + j]: column j belongs to lane j
  __shared__ __align__(16) double aring[OQ ? OQ : 1][LQ];     // alpha vectors of the pending outer products
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPS, j = lane % LPS;
  const int K = p.K, T = p.T;
  const long long seq = (long long)blockIdx.x * SPW + sub;
  const long long b = seq < p.B ? seq : (long long)p.B - 1;
  const bool valid = seq < p.B && (!p.active || p.active[b] != 0);
  const bool act = j < K;
  const bool fwd_only = (p.flags & 1) != 0;
  const bool store = act && valid && !fwd_only;

  // global shift of log A so that P = exp(log A - gmax) <= 1
  float gmf = -INFINITY;
  for (int e = j; e < K * K; e += LPS) gmf = fmaxf(gmf, (float)p.log_A[e]);
  const double gmax = (double)tile_maxf<LPS>(gmf);

  const double* node = p.node + (size_t)b * T * K;
  double* Ez = p.Ez + (size_t)b * T * K;
  const double NEG = -1.0e300;

  // ------------------------------------ forward ---------------------------------------------
  double logZ, cprod = 1.0, e, c, invc;
  double x_last;   // node[T-1][j], reused by the backward sweep
  {
    double Pc[LPS];   // column j of P:  Pc[i] = P[i][j]
#pragma unroll
    for (int i = 0; i < LPS; ++i) Pc[i] = (act && i < K) ? exp(p.log_A[i * K + j] - gmax) : 0.0;

    const double x0 = act ? node[j] : NEG;
    x_last = x0;
    const double a0 = act ? p.log_pi[j] + x0 : NEG;
    float mf = tile_maxf<LPS>((float)a0);
    e = act ? exp(a0 - (double)mf) : 0.0;
    logZ = (double)mf;
    double x_nxt = (act && T > 1) ? node[K + j] : NEG;
    float mfn = tile_maxf<LPS>((float)x_nxt);
    double exn = act ? exp(x_nxt - (double)mfn) : 0.0;
    // node potentials are prefetched three steps ahead (a step is shorter than one DRAM round trip)
    double xq1 = (act && T > 2) ? node[(size_t)2 * K + j] : NEG;
    double xq2 = (act && T > 3) ? node[(size_t)3 * K + j] : NEG;
    for (int t = 1; t < T; ++t) {
      double* buf = sbuf[sub][t & 1];
      buf[j] = e;
      const double x_pre = xq1;                                                          // node[t+1]
      xq1 = xq2;
      xq2 = (act && t + 3 < T) ? node[(size_t)(t + 3) * K + j] : NEG;
      __syncwarp();
      double pa[4] = {0.0, 0.0, 0.0, 0.0}, sa[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int i2 = 0; i2 < LPS / 2; ++i2) {
        const double2 a2 = *reinterpret_cast<const double2*>(&buf[2 * i2]);
        pa[(2 * i2) & 3] = fma(a2.x, Pc[2 * i2], pa[(2 * i2) & 3]);
        pa[(2 * i2 + 1) & 3] = fma(a2.y, Pc[2 * i2 + 1], pa[(2 * i2 + 1) & 3]);
        sa[(2 * i2) & 3] += a2.x;
        sa[(2 * i2 + 1) & 3] += a2.y;
      }
      const double pj = (pa[0] + pa[1]) + (pa[2] + pa[3]);
      c = (sa[0] + sa[1]) + (sa[2] + sa[3]);          // sum of the previous message, same in every lane
      invc = hmm_rcp(c);
      if (store) Ez[(size_t)(t - 1) * K + j] = e * invc;          // alpha_{t-1}, normalised
      logZ += (double)mfn;
      cprod *= c;
      if ((t & 7) == 7) {
        logZ += log(cprod);
        cprod = 1.0;
      }
      e = pj * invc * exn;                                       // message t (sums to c_t)
      x_last = x_nxt;
      x_nxt = x_pre;
      mfn = tile_maxf<LPS>((float)x_nxt);
      exn = act ? exp(x_nxt - (double)mfn) : 0.0;
    }
    c = tile_sum<LPS>(e);
    invc = 1.0 / c;
    cprod *= c;
    logZ += log(cprod) + (double)(T - 1) * gmax;
    if (j == 0 && valid && p.logZ) p.logZ[b] = logZ;
    if (fwd_only) return;
    if (store) Ez[(size_t)(T - 1) * K + j] = e * invc;            // alpha_{T-1} = E[z_{T-1}]
  }

  // ------------------------------------ backward --------------------------------------------
  double Pr[LPS];   // row j of P:  Pr[i] = P[j][i]
#pragma unroll
  for (int i = 0; i < LPS; ++i) Pr[i] = (act && i < K) ? exp(p.log_A[j * K + i] - gmax) : 0.0;
  double O[OQ ? 1 : LPS];    // column j of sum_t alpha_t (x) w_{t+1} / s_t
  double wq[OQ ? OQ : 1];    // OQ: the scaled w_{t+1}[j] of the pending outer products
  int npend = 0;
  if (OQ) {
#pragma unroll
    for (int i = 0; i < 32; ++i) Osm[i * 32 + j] = 0.0;
#pragma unroll
    for (int q = 0; q < (OQ ? OQ : 1); ++q) { wq[q] = 0.0; aring[q][j] = 0.0; }
  } else {
#pragma unroll
    for (int i = 0; i < (OQ ? 1 : LPS); ++i) O[i] = 0.0;
  }
  // OQ: Osm[:, j] += sum_q aring[q][:] * wq[q]   (own column: no other lane touches it)
  auto flush = [&]() {
#pragma unroll
    for (int i2 = 0; i2 < 16; ++i2) {
      double o0 = Osm[(2 * i2) * 32 + j], o1 = Osm[(2 * i2 + 1) * 32 + j];
#pragma unroll
      for (int q = 0; q < (OQ ? OQ : 1); ++q) {
        const double2 a2 = *reinterpret_cast<const double2*>(&aring[q][2 * i2]);
        o0 = fma(a2.x, wq[q], o0);
        o1 = fma(a2.y, wq[q], o1);
      }
      Osm[(2 * i2) * 32 + j] = o0;
      Osm[(2 * i2 + 1) * 32 + j] = o1;
    }
#pragma unroll
    for (int q = 0; q < (OQ ? OQ : 1); ++q) wq[q] = 0.0;
  };
  __syncwarp();
  double bet = act ? 1.0 : 0.0;     // beta_{t+1}, arbitrary positive scale
  double kb = 1.0;                  // lagged rescaling of beta
  double exl;                       // exp(node_{t+1} - max)
  {
    const float mfl = tile_maxf<LPS>((float)x_last);
    exl = act ? exp(x_last - (double)mfl) : 0.0;
  }
  double at = (act && T > 1) ? Ez[(size_t)(T - 2) * K + j] : 0.0;     // alpha_t (same thread wrote it)
  double x_pre = (act && T > 1) ? node[(size_t)(T - 2) * K + j] : NEG;
  // three-deep prefetch queues for alpha_{t-1}, node[t-1]
  double aq1 = (act && T > 2) ? Ez[(size_t)(T - 3) * K + j] : 0.0, aq2 = (act && T > 3) ? Ez[(size_t)(T - 4) * K + j] : 0.0;
  double nq1 = (act && T > 2) ? node[(size_t)(T - 3) * K + j] : NEG, nq2 = (act && T > 3) ? node[(size_t)(T - 4) * K + j] : NEG;
  double prd_prev = 0.0, at_prev = 0.0, w_prev = 0.0;                 // finished one step late
  for (int t = T - 2; t >= 0; --t) {
    const int sl = 2 + 3 * (t & 1);
    double* wb = sbuf[sub][sl];
    double* pbf = sbuf[sub][sl + 1];
    double* ab = OQ ? aring[npend] : sbuf[sub][sl + 2];
    const double w = exl * bet * kb;          // w_{t+1}[j]
    wb[j] = w;
    pbf[j] = prd_prev;                        // alpha_{t+1} .* (P w_{t+2})
    ab[j] = at_prev;                          // alpha_{t+1}
    const double at_next = aq1, x_next = nq1;                                       // alpha_{t-1}, node[t-1]
    aq1 = aq2;
    nq1 = nq2;
    aq2 = (act && t > 2) ? Ez[(size_t)(t - 3) * K + j] : 0.0;
    nq2 = (act && t > 2) ? node[(size_t)(t - 3) * K + j] : NEG;
    __syncwarp();
    double pb[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i2 = 0; i2 < LPS / 2; ++i2) {
      const double2 w2 = *reinterpret_cast<const double2*>(&wb[2 * i2]);
      const double2 p2 = *reinterpret_cast<const double2*>(&pbf[2 * i2]);
      pb[(2 * i2) & 3] = fma(w2.x, Pr[2 * i2], pb[(2 * i2) & 3]);
      pb[(2 * i2 + 1) & 3] = fma(w2.y, Pr[2 * i2 + 1], pb[(2 * i2 + 1) & 3]);
      ss[(2 * i2) & 3] += p2.x;
      ss[(2 * i2 + 1) & 3] += p2.y;
    }
    const double bu = (pb[0] + pb[1]) + (pb[2] + pb[3]);       // (P w_{t+1})[j]
    if (t < T - 2) {
      // finish step t+1: s_{t+1} = <alpha_{t+1}, P w_{t+2}>
      const double invs = hmm_rcp((ss[0] + ss[1]) + (ss[2] + ss[3]));
      if (act && valid) Ez[(size_t)(t + 1) * K + j] = prd_prev * invs;
      const double wsc = w_prev * invs;
      if (OQ) {
#pragma unroll
        for (int q = 0; q < (OQ ? OQ : 1); ++q)
          if (q == npend) wq[q] = wsc;
        if (++npend == OQ) {          // warp-uniform
          flush();
          npend = 0;
          __syncwarp();               // the ring is rewritten from the next iteration on
        }
      } else {
#pragma unroll
        for (int i2 = 0; i2 < LPS / 2; ++i2) {
          const double2 a2 = *reinterpret_cast<const double2*>(&ab[2 * i2]);
          O[2 * i2] = fma(a2.x, wsc, O[2 * i2]);
          O[2 * i2 + 1] = fma(a2.y, wsc, O[2 * i2 + 1]);
        }
      }
      kb = invs;
    }
    prd_prev = at * bu;
    at_prev = at;
    w_prev = w;
    bet = bu;
    at = at_next;
    {
      // exp factor of node[t] (prefetched one iteration ago) for the next iteration
      const double xv = act ? x_pre : NEG;
      const float mfl = tile_maxf<LPS>((float)xv);
      exl = act ? exp(xv - (double)mfl) : 0.0;
    }
    x_pre = x_next;
  }
  double ez0 = e * invc;     // T == 1: alpha_0
  if (T > 1) {
    // finish step 0
    const double s0 = tile_sum<LPS>(prd_prev);
    const double invs = 1.0 / s0;
    ez0 = prd_prev * invs;
    if (act && valid) Ez[j] = ez0;
    const double wsc = w_prev * invs;
    if (OQ) {
      __syncwarp();
      aring[npend][j] = at_prev;
#pragma unroll
      for (int q = 0; q < (OQ ? OQ : 1); ++q)
        if (q == npend) wq[q] = wsc;
      __syncwarp();
      flush();                        // the slots past npend carry wq = 0 and finite (stale or zero) alpha values
    } else {
      double* ab = sbuf[sub][0];
      ab[j] = at_prev;
      __syncwarp();
#pragma unroll
      for (int i2 = 0; i2 < LPS / 2; ++i2) {
        const double2 a2 = *reinterpret_cast<const double2*>(&ab[2 * i2]);
        O[2 * i2] = fma(a2.x, wsc, O[2 * i2]);
        O[2 * i2 + 1] = fma(a2.y, wsc, O[2 * i2 + 1]);
      }
    }
  }
  __syncwarp();
  if (act && valid) {
    if (p.Ez0) p.Ez0[(size_t)b * K + j] = ez0;
    if (p.Etrans) {
#pragma unroll
      for (int i = 0; i < LPS; ++i)
        if (i < K) p.Etrans[(size_t)b * K * K + i * K + j] = exp(p.log_A[i * K + j] - gmax) * (OQ ? Osm[i * 32 + j] : O[OQ ? 0 : i]);
    }
  }
}

// Viterbi: max-product forward with byte back-pointers in the workspace, then a serial back-track.
struct ViterbiParams {
  int B, T, K;
  const double *log_pi, *log_A, *node;
  int32_t* path;           // (B,T)
  uint8_t* back;           // workspace (B,T,K)
};

template <int LPS>
__global__ void __launch_bounds__(32) hmm_viterbi_kernel(const ViterbiParams p) {
  constexpr int SPW = 32 / LPS;
  constexpr int LP = (LPS + 1) & ~1;
  __shared__ __align__(16) double sbuf[SPW][LP < 2 ? 2 : LP];
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPS, j = lane % LPS;
  const int K = p.K, T = p.T;
  const long long seq = (long long)blockIdx.x * SPW + sub;
  const bool valid = seq < p.B;
  const long long b = valid ? seq : (long long)p.B - 1;
  const bool act = j < K;
  double* dbuf = sbuf[sub];
  const double NEG = -1.0e300;
  double Ac[LPS];   // column j of log A
#pragma unroll
  for (int i = 0; i < LPS; ++i) Ac[i] = (act && i < K) ? p.log_A[i * K + j] : NEG;
  const double* node = p.node + (size_t)b * T * K;
  uint8_t* back = p.back + (size_t)b * T * K;
  double delta = act ? p.log_pi[j] + node[j] : NEG;
  for (int t = 1; t < T; ++t) {
    dbuf[j] = delta;
    __syncwarp();
    double best = NEG * 2.0;
    int arg = 0;
#pragma unroll
    for (int i = 0; i < LPS; ++i) {
      const double v = dbuf[i] + Ac[i];
      if (v > best) { best = v; arg = i; }      // strict '>' keeps the FIRST maximal index
    }
    delta = act ? best + node[(size_t)t * K + j] : NEG;
    if (act && valid) back[(size_t)t * K + j] = (uint8_t)arg;
    __syncwarp();
  }
  // argmax over the tile (first max): reduce (value, index) pairs
  double bv = delta;
  int bi = j;
#pragma unroll
  for (int o = LPS / 2; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o, LPS);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o, LPS);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  __syncwarp();
  __threadfence_block();
  if (j == 0 && valid) {
    int32_t* path = p.path + (size_t)b * T;
    int z = bi;
    path[T - 1] = z;
    for (int t = T - 2; t >= 0; --t) {
      z = back[(size_t)(t + 1) * K + z];
      path[t] = z;
    }
  }
}

template <int LPS>
static cudaError_t launch_hmm(const HmmParams& p, cudaStream_t s) {
  constexpr int SPW = 32 / LPS;
  if (LPS == 32 && SVAE_HMM_OSMEM) {
    // 16 one-warp CTAs of 11 KB each per SM: ask for the large shared-memory carve-out (a hint; per device, cheap)
    cudaFuncSetAttribute(hmm_estep_kernel<LPS>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  }
  hmm_estep_kernel<LPS><<<(unsigned)((p.B + SPW - 1) / SPW), 32, 0, s>>>(p);
  return cudaGetLastError();
}
template <int LPS>
static cudaError_t launch_vit(const ViterbiParams& p, cudaStream_t s) {
  constexpr int SPW = 32 / LPS;
  hmm_viterbi_kernel<LPS><<<(unsigned)((p.B + SPW - 1) / SPW), 32, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t hmm_estep_launch(const HmmParams& p, cudaStream_t s) {
  if (p.K <= 2) return launch_hmm<2>(p, s);
  if (p.K <= 4) return launch_hmm<4>(p, s);
  if (p.K <= 8) return launch_hmm<8>(p, s);
  if (p.K <= 16) return launch_hmm<16>(p, s);
  return launch_hmm<32>(p, s);
}
cudaError_t hmm_viterbi_launch(const ViterbiParams& p, cudaStream_t s) {
  if (p.K <= 2) return launch_vit<2>(p, s);
  if (p.K <= 4) return launch_vit<4>(p, s);
  if (p.K <= 8) return launch_vit<8>(p, s);
  if (p.K <= 16) return launch_vit<16>(p, s);
  return launch_vit<32>(p, s);
}

}  // namespace svae

extern std::atomic<uint64_t> g_svae_launches;

extern "C" int svae_hmm_estep_masked(int B, int T, int K, int flags, const double* log_pi, const double* log_A,
                                     const double* node, const int* active, double* logZ, double* Ez0,
                                     double* Etrans, double* Ez, void* stream) {
  if (B <= 0 || T <= 0 || K <= 0) return SVAE_ERR_BAD_ARG;
  if (!log_pi || !log_A || !node) return SVAE_ERR_BAD_ARG;
  const bool fwd_only = (flags & SVAE_HMM_LOGZ_ONLY) != 0;
  if (!fwd_only && !Ez) return SVAE_ERR_BAD_ARG;
  if (K > SVAE_HMM_MAX_K) return SVAE_ERR_UNSUPPORTED;
  svae::HmmParams p;
  p.B = B; p.T = T; p.K = K; p.flags = fwd_only ? 1 : 0;
  p.log_pi = log_pi; p.log_A = log_A; p.node = node;
  p.logZ = logZ; p.Ez0 = Ez0; p.Etrans = Etrans; p.Ez = Ez;
  p.active = active;
  cudaError_t e = svae::hmm_estep_launch(p, static_cast<cudaStream_t>(stream));
  g_svae_launches.fetch_add(1);
  return e == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}

extern "C" int svae_hmm_estep(int B, int T, int K, int flags, const double* log_pi, const double* log_A,
                              const double* node, double* logZ, double* Ez0, double* Etrans, double* Ez,
                              void* stream) {
  return svae_hmm_estep_masked(B, T, K, flags, log_pi, log_A, node, nullptr, logZ, Ez0, Etrans, Ez, stream);
}

int svae_hmm_viterbi_wide_launch(int B, int T, int K, const double* log_pi, const double* log_A, const double* node,
                                 int32_t* path, void* workspace, cudaStream_t s);   // hmm_wide.cu

extern "C" size_t svae_hmm_viterbi_workspace_bytes(int B, int T, int K) {
  if (B <= 0 || T <= 0 || K <= 0) return 0;
  return (size_t)B * (size_t)T * (size_t)K * (K > SVAE_HMM_MAX_K ? 2 : 1);    // 8- or 16-bit back-pointers
}

extern "C" int svae_hmm_viterbi(int B, int T, int K, const double* log_pi, const double* log_A, const double* node,
                                int32_t* path, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || T <= 0 || K <= 0) return SVAE_ERR_BAD_ARG;
  if (!log_pi || !log_A || !node || !path || !workspace) return SVAE_ERR_BAD_ARG;
  if (K > SVAE_HMM_WIDE_MAX_K) return SVAE_ERR_UNSUPPORTED;
  if (workspace_bytes < svae_hmm_viterbi_workspace_bytes(B, T, K)) return SVAE_ERR_WORKSPACE;
  if (K > SVAE_HMM_MAX_K)
    return svae_hmm_viterbi_wide_launch(B, T, K, log_pi, log_A, node, path, workspace, static_cast<cudaStream_t>(stream));
  svae::ViterbiParams p;
  p.B = B; p.T = T; p.K = K; p.log_pi = log_pi; p.log_A = log_A; p.node = node;
  p.path = path; p.back = static_cast<uint8_t*>(workspace);
  cudaError_t e = svae::hmm_viterbi_launch(p, static_cast<cudaStream_t>(stream));
  g_svae_launches.fetch_add(1);
  return e == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}
