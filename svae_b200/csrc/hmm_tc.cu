// hmm_tc.cu -- HMM local E-step for 17..32 states on the fp64 tensor-core instruction (mma.m8n8k4), sm_100a.
//
// What it replaces in the reference is what hmm_estep.cu replaces: svae/hmm/cython_hmm_inference.pyx:93-166 (hmm_logZ /
// hmm_logZ_grad: with cotangent 1 the E-step) and svae/hmm/hmm_inference.py:21-41.
//
// Why.  ncu on hmm_estep_kernel<32> (profiles/r02_hmm_*_summary.txt): one warp per sequence, the K x K matrix-vector products
// as 32 DFMA per lane against a message broadcast through shared memory -- 146 shared-memory wavefronts and ~500
// warp-instructions per sequence-step; short-scoreboard (shared-memory) stalls dominate, and neither more resident warps
// (the -DSVAE_HMM_OSMEM build: 16 instead of 8 per SM, same 1.13 ms) nor fewer registers move it.  The dense contraction
// belongs on the tensor cores, with the BATCH as the M dimension:
//   * 8 sequences form the rows of every fragment (row g = lane >> 2); a step's 8 x 32 by 32 x 32 product is 32 mma.m8n8k4
//     whose B operands -- P = exp(log A - max) forward, P' backward -- stay in registers for the whole sweep.  The default
//     sweep (hmm_tc_sweep4_kernel) spreads the four n-tiles of those 8 sequences over the four warps of a CTA; the one-warp
//     sweep (hmm_tc_sweep_kernel, SVAE_HMM_TC_WARPS=1) keeps all of it in one warp and needs no shared memory at all;
//   * state vectors live in the ACCUMULATOR layout (lane (g,q): states 8 nt + 2 q + e, nt < 4, e < 2), and the A operand of
//     slab s is defined as element [s >> 1][s & 1] of it, i.e. the contraction runs over the states in the order
//     8 (s >> 1) + 2 q + (s & 1): the output of one step is the input of the next with NO shuffle or shared memory (the P
//     fragment tables absorb the relabelling; scripts/dev_hmm_tc_numpy.py pins the index maps);
//   * row sums / maxima are 8 local operations + 2 quad shuffles; node potentials are read as 64-byte quad segments.
//   * the per-sequence transition counts  E[#i->j] = P[i][j] sum_t ahat_t[i] w_{t+1}[j] / s_t  do not fit next to 8 sequences
//     per warp (8 x 32 x 32 accumulators), so the sweep leaves ahat_t and w_{t+1}/s_t in a workspace and a second kernel
//     forms the 32 x 32 x T product of each sequence with the contraction over TIME on the tensor cores (16 mma per 4 steps).
// Same scaled recursions and outputs as hmm_estep.cu (the tests hold both to the same golden vectors).
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <cstdlib>
#include "svae_b200.h"

extern std::atomic<uint64_t> g_svae_launches;

namespace svae {

__device__ __forceinline__ void hmm_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_maxf(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ double tc_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// -DSVAE_HMM_TC_FASTEXP=1: exp() for the arguments of the sweeps (node - max: <= 0 up to the float rounding of the maximum).
// Same range reduction as the library routine, degree-13 Taylor polynomial evaluated by Estrin's scheme (5 dependent
// operations instead of a 13-deep Horner chain: the four-warp sweep is bound by dependent-issue latency), 0 below -700, no
// overflow / denormal handling; <= 4 ulp against std::exp over [-705, 1] (host test of the same source: 2e7 random points).
#ifndef SVAE_HMM_TC_FASTEXP
#define SVAE_HMM_TC_FASTEXP 0
#endif
__device__ __forceinline__ double tc_exp(double x) {
#if SVAE_HMM_TC_FASTEXP
  const double L2E = 1.4426950408889634074, LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  const double SHIFT = 6755399441055744.0;            // 1.5 * 2^52: adding it leaves round(x log2 e) in the low mantissa bits
  if (x < -700.0) return 0.0;
  const double kd = fma(x, L2E, SHIFT);
  const int ki = __double2loint(kd);
  const double k = kd - SHIFT;
  double r = fma(-k, LN2_HI, x);
  r = fma(-k, LN2_LO, r);
  const double c2 = 1.0 / 2, c3 = 1.0 / 6, c4 = 1.0 / 24, c5 = 1.0 / 120, c6 = 1.0 / 720, c7 = 1.0 / 5040, c8 = 1.0 / 40320,
               c9 = 1.0 / 362880, c10 = 1.0 / 3628800, c11 = 1.0 / 39916800, c12 = 1.0 / 479001600, c13 = 1.0 / 6227020800.0;
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double p01 = 1.0 + r, p23 = fma(c3, r, c2), p45 = fma(c5, r, c4), p67 = fma(c7, r, c6), p89 = fma(c9, r, c8),
               pab = fma(c11, r, c10), pcd = fma(c13, r, c12);
  const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45), q2 = fma(pab, r2, p89);
  const double s0 = fma(q1, r4, q0), s1 = fma(pcd, r4, q2);
  const double pp = fma(s1, r8, s0);
  return pp * __hiloint2double((ki + 1023) << 20, 0);   // 2^k, k >= -1010 here
#else
  return exp(x);
#endif
}

struct HmmTcParams {
  int B, T, K, fwd_only;
  const double *log_pi, *log_A, *node;   // (K), (K,K) row = from, (B,T,K)
  const int* active;
  double *logZ, *Ez0, *Ez;               // (B), (B,K), (B,T,K)
  double *wsA, *wsW;                     // workspace (B,T,K) each: ahat_t; w_{t+1} / s_t (rows 0 .. T-2)
};

// max of log_A over the K x K entries, the same value in every lane
__device__ __forceinline__ double tc_gmax(const double* log_A, int K, int lane) {
  float m = -INFINITY;
  for (int e = lane; e < K * K; e += 32) m = fmaxf(m, (float)log_A[e]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  return (double)m;
}

// ---- forward-backward sweep: one warp per 8 sequences ------------------------------------------------------------
__global__ void __launch_bounds__(32) hmm_tc_sweep_kernel(const HmmTcParams p) {
  const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
  const int K = p.K, T = p.T;
  const long long seq = (long long)blockIdx.x * 8 + g;
  const long long b = seq < p.B ? seq : (long long)p.B - 1;
  const bool valid = seq < p.B && (!p.active || p.active[b] != 0);
  const double NEG = -INFINITY;
  const double gmax = tc_gmax(p.log_A, K, lane);

  // this lane's 8 states: st(nt, e) = 8 nt + 2 q + e
  bool in[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) in[nt][e] = (8 * nt + 2 * q + e) < K;

  const double* node = p.node + (size_t)b * T * K + 2 * q;
  double* Ez = p.Ez ? p.Ez + (size_t)b * T * K + 2 * q : nullptr;
  double* wsA = p.wsA ? p.wsA + (size_t)b * T * K + 2 * q : nullptr;
  double* wsW = p.wsW ? p.wsW + (size_t)b * T * K + 2 * q : nullptr;

  auto load_row = [&](const double* base, int t, double (&x)[4][2], double fill) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) x[nt][e] = in[nt][e] ? __ldg(base + (size_t)t * K + 8 * nt + e) : fill;
  };
  auto store_row = [&](double* base, int t, const double (&x)[4][2]) {
    if (!valid) return;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (in[nt][e]) base[(size_t)t * K + 8 * nt + e] = x[nt][e];
  };
  auto row_maxf = [&](const double (&x)[4][2]) {
    float m = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) m = fmaxf(m, fmaxf((float)x[nt][0], (float)x[nt][1]));
    return quad_maxf(m);
  };
  auto row_sum = [&](const double (&x)[4][2]) {
    const double s = ((x[0][0] + x[0][1]) + (x[1][0] + x[1][1])) + ((x[2][0] + x[2][1]) + (x[3][0] + x[3][1]));
    return quad_sum(s);
  };
  // out = v (8 x 32, this warp's 8 sequences) times the 32 x 32 matrix whose fragments are F[s][nt]
  auto matvec = [&](const double (&v)[4][2], const double (&F)[8][4], double (&out)[4][2]) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) out[nt][0] = out[nt][1] = 0.0;
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) hmm_dmma(out[nt][0], out[nt][1], v[s >> 1][s & 1], F[s][nt]);
  };

  // ------------------------------------------------ forward ------------------------------------------------
  double e_[4][2];          // message (unnormalised)
  double logZ;
  {
    double F[8][4];         // F[s][nt] = P[8 (s>>1) + 2 q + (s&1)][8 nt + g]
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int i = 8 * (s >> 1) + 2 * q + (s & 1), j = 8 * nt + g;
        F[s][nt] = (i < K && j < K) ? exp(p.log_A[i * K + j] - gmax) : 0.0;
      }
    double x[4][2], xn1[4][2], xn2[4][2];
    load_row(node, 0, x, NEG);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) x[nt][e] = in[nt][e] ? p.log_pi[8 * nt + 2 * q + e] + x[nt][e] : NEG;
    {
      const double m = (double)row_maxf(x);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) e_[nt][e] = in[nt][e] ? exp(x[nt][e] - m) : 0.0;
      logZ = m;
    }
    if (T > 1) load_row(node, 1, xn1, NEG);
    if (T > 2) load_row(node, 2, xn2, NEG);
    double cprod = 1.0;
    for (int t = 1; t < T; ++t) {
      // node potentials of step t are in xn1; fetch t + 2 while this step computes
      double xq[4][2];
      if (t + 2 < T) load_row(node, t + 2, xq, NEG);
      double d[4][2];
      matvec(e_, F, d);
      const double c = row_sum(e_);
      const double invc = tc_rcp(c);
      const double m = (double)row_maxf(xn1);
      double ah[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          ah[nt][e] = e_[nt][e] * invc;                                   // ahat_{t-1}
          e_[nt][e] = in[nt][e] ? d[nt][e] * invc * exp(xn1[nt][e] - m) : 0.0;
        }
      if (!p.fwd_only) store_row(wsA, t - 1, ah);
      logZ += m;
      cprod *= c;
      if ((t & 7) == 7) {
        logZ += log(cprod);
        cprod = 1.0;
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          xn1[nt][e] = xn2[nt][e];
          xn2[nt][e] = xq[nt][e];
        }
    }
    const double c = row_sum(e_);
    const double invc = 1.0 / c;
    cprod *= c;
    logZ += log(cprod) + (double)(T - 1) * gmax;
    if (q == 0 && valid && p.logZ) p.logZ[b] = logZ;
    if (p.fwd_only) return;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) e_[nt][e] *= invc;                      // ahat_{T-1} = E[z_{T-1}]
    store_row(Ez, T - 1, e_);
    if (T == 1 && p.Ez0) store_row(p.Ez0 + (size_t)b * K + 2 * q, 0, e_);
  }
  if (T == 1) return;

  // ------------------------------------------------ backward -----------------------------------------------
  double F[8][4];           // F[s][nt] = P'[k][j] = P[8 nt + g][8 (s>>1) + 2 q + (s&1)]
#pragma unroll
  for (int s = 0; s < 8; ++s)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int k = 8 * (s >> 1) + 2 * q + (s & 1), j = 8 * nt + g;
      F[s][nt] = (k < K && j < K) ? exp(p.log_A[j * K + k] - gmax) : 0.0;
    }
  __syncwarp();             // the ahat rows this warp stored are read back below (same lanes, same addresses)
  double bet[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) bet[nt][e] = in[nt][e] ? 1.0 : 0.0;
  double xn[4][2], an[4][2];          // node[t+1], ahat_t of the CURRENT iteration
  load_row(node, T - 1, xn, NEG);
  {
    // ahat_{T-2} was stored by this very lane; re-read through the normal path (not __ldg: written in this kernel)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) an[nt][e] = (in[nt][e] && valid) ? wsA[(size_t)(T - 2) * K + 8 * nt + e] : (in[nt][e] ? 1.0 / K : 0.0);
  }
  for (int t = T - 2; t >= 0; --t) {
    double xp[4][2], ap[4][2];          // prefetch for iteration t - 1: node[t], ahat_{t-1}
    if (t > 0) {
      load_row(node, t, xp, NEG);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) ap[nt][e] = (in[nt][e] && valid) ? wsA[(size_t)(t - 1) * K + 8 * nt + e] : (in[nt][e] ? 1.0 / K : 0.0);
    }
    const double m = (double)row_maxf(xn);
    double w[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) w[nt][e] = in[nt][e] ? exp(xn[nt][e] - m) * bet[nt][e] : 0.0;      // w_{t+1}
    double bu[4][2];
    matvec(w, F, bu);                                                     // P w_{t+1}
    double prd[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) prd[nt][e] = an[nt][e] * bu[nt][e];
    const double s = row_sum(prd);
    const double invs = tc_rcp(s);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        prd[nt][e] *= invs;                                               // E[z_t]
        w[nt][e] *= invs;                                                 // w_{t+1} / s_t
        bet[nt][e] = bu[nt][e] * invs;
      }
    store_row(Ez, t, prd);
    store_row(wsW, t, w);
    if (t == 0 && p.Ez0) store_row(p.Ez0 + (size_t)b * K + 2 * q, 0, prd);
    if (t > 0) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          xn[nt][e] = xp[nt][e];
          an[nt][e] = ap[nt][e];
        }
    }
  }
}

// ---- forward-backward sweep, FOUR warps per 8 sequences -----------------------------------------------------------
// ncu on the one-warp sweep above at the BASELINE batch (2048 sequences = 256 warps on 592 schedulers): 2 886 cycles per
// step for a lone warp -- its 32 dependent-chain mma run on ONE scheduler's quarter of the fp64 tensor pipe and its 8 exp()
// per lane sit behind them (stall reason "wait" 60 %).  Here warp w of a 4-warp CTA owns n-tile w of the same 8 sequences
// (states 8 w + 2 q + e): 8 mma (two accumulator chains) and 2 exp() per lane and step; the full state vector the mma
// need as A operands is exchanged through 4 KB of shared memory with ONE barrier per step, which also carries the
// per-warp partial row sums / maxima.  To get by with one barrier every row normaliser is applied one step late:
//   forward : e_t = (e_{t-1} P) / sum(e_{t-1}) .* exp(node_t - max_t)         (partials of sum(e_{t-1}) travel with e_{t-1})
//   backward: U_{t+1} = exp(node_{t+1} - max) .* B_{t+1} is published BEFORE s_{t+1} = <ahat_{t+1}, B_{t+1}> is known;
//             B_t = P U_{t+1} / s_{t+1};  E[z_{t+1}] and W_{t+1} = U_{t+2} / (s_{t+2} s_{t+1}) are stored one step late.
// scripts/dev_hmm_tc_numpy.py (sweep_cta4) is this kernel statement by statement in NumPy, against the oracle.
//
// Build-time knobs, each measured on a B200 (profiles/r02_hmm_tc_summary.txt) -- the defaults are the winners:
// prefetch depth (steps) of the global loads, forward / backward sweep.  Measured (profiles/r02_hmm_tc_summary.txt): deeper
// rings cost registers (80 -> 110) and occupancy at large batches for no gain at the BASELINE batch.
#ifndef SVAE_HMM_TC_PF
#define SVAE_HMM_TC_PF 2
#endif
// (backward depth: a template parameter chosen at launch -- 2 while every CTA of the grid is resident anyway, 1 beyond)
#ifndef SVAE_HMM_TC_TRANS_AHEAD
#define SVAE_HMM_TC_TRANS_AHEAD 1     // groups of 4 time steps in flight ahead of the one being multiplied (1 or 2)
#endif
#ifndef SVAE_HMM_TC_CHAINS
#define SVAE_HMM_TC_CHAINS 2          // accumulate chains of a tile's 8 mma: 2 (x4 deep), 4 (x2 deep) or 8 (independent)
#endif
template <int PB>
__global__ void __launch_bounds__(128) hmm_tc_sweep4_kernel(const HmmTcParams p) {
  constexpr int PF = SVAE_HMM_TC_PF;
  __shared__ __align__(16) double vec[2][4][2][32];      // [buffer][n-tile][element][lane]
  __shared__ double psum[2][4][8];                       // [buffer][warp][row]: partial row sums
  __shared__ float pmax[2][4][8];                        // partial row maxima of the next node potentials
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  const int K = p.K, T = p.T;
  const long long seq = (long long)blockIdx.x * 8 + g;
  const long long b = seq < p.B ? seq : (long long)p.B - 1;
  const bool valid = seq < p.B && (!p.active || p.active[b] != 0);
  const double NEG = -INFINITY;
  const double gmax = tc_gmax(p.log_A, K, lane);
  const int s0 = 8 * w + 2 * q;                           // this lane's two states: s0, s0 + 1
  const bool in0 = s0 < K, in1 = s0 + 1 < K;

  const double* node = p.node + (size_t)b * T * K + s0;
  double* Ez = p.Ez ? p.Ez + (size_t)b * T * K + s0 : nullptr;
  double* wsA = p.wsA ? p.wsA + (size_t)b * T * K + s0 : nullptr;
  double* wsW = p.wsW ? p.wsW + (size_t)b * T * K + s0 : nullptr;

  auto ld2n = [&](int t, double& x0, double& x1) {       // node potentials of this lane's two states at time t
    x0 = in0 ? __ldg(node + (size_t)t * K) : NEG;
    x1 = in1 ? __ldg(node + (size_t)t * K + 1) : NEG;
  };
  auto st2 = [&](double* base, int t, double v0, double v1) {
    if (!valid) return;
    if (in0) base[(size_t)t * K] = v0;
    if (in1) base[(size_t)t * K + 1] = v1;
  };
  auto pmaxf2 = [&](double x0, double x1) { return quad_maxf(fmaxf((float)x0, (float)x1)); };
  // own n-tile of (vector in vec[buf]) times the matrix whose fragments are F[s]: two accumulator chains
  auto tile_mv = [&](int buf, const double (&F)[8], double& d0, double& d1) {
#if SVAE_HMM_TC_CHAINS == 8
    // eight independent mma (no accumulate chain at all), then an add tree
    double c[8][2];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      c[s][0] = c[s][1] = 0.0;
      hmm_dmma(c[s][0], c[s][1], vec[buf][s >> 1][s & 1][lane], F[s]);
    }
    d0 = ((c[0][0] + c[1][0]) + (c[2][0] + c[3][0])) + ((c[4][0] + c[5][0]) + (c[6][0] + c[7][0]));
    d1 = ((c[0][1] + c[1][1]) + (c[2][1] + c[3][1])) + ((c[4][1] + c[5][1]) + (c[6][1] + c[7][1]));
#elif SVAE_HMM_TC_CHAINS == 4
    double c[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k][0] = c[k][1] = 0.0;
#pragma unroll
    for (int s = 0; s < 8; ++s) hmm_dmma(c[s & 3][0], c[s & 3][1], vec[buf][s >> 1][s & 1][lane], F[s]);
    d0 = (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
    d1 = (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
#else
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
    for (int s = 0; s < 8; s += 2) {
      hmm_dmma(c0, c1, vec[buf][s >> 1][0][lane], F[s]);
      hmm_dmma(c2, c3, vec[buf][s >> 1][1][lane], F[s + 1]);
    }
    d0 = c0 + c2;
    d1 = c1 + c3;
#endif
  };
  auto sum4 = [&](int buf) { return (psum[buf][0][g] + psum[buf][1][g]) + (psum[buf][2][g] + psum[buf][3][g]); };
  auto max4 = [&](int buf) { return (double)fmaxf(fmaxf(pmax[buf][0][g], pmax[buf][1][g]), fmaxf(pmax[buf][2][g], pmax[buf][3][g])); };

  // ------------------------------------------------ forward ------------------------------------------------
  double e0, e1;            // this lane's two entries of the (unnormalised) message
  {
    double F[8];            // F[s] = P[8 (s>>1) + 2 q + (s&1)][8 w + g]
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int i = 8 * (s >> 1) + 2 * q + (s & 1), j = 8 * w + g;
      F[s] = (i < K && j < K) ? exp(p.log_A[i * K + j] - gmax) : 0.0;
    }
    double x0, x1;
    ld2n(0, x0, x1);
    x0 = in0 ? p.log_pi[s0] + x0 : NEG;
    x1 = in1 ? p.log_pi[s0 + 1] + x1 : NEG;
    {
      const float pm = pmaxf2(x0, x1);
      if (q == 0) pmax[0][w][g] = pm;
    }
    __syncthreads();
    double logZ = max4(0);
    e0 = in0 ? tc_exp(x0 - logZ) : 0.0;
    e1 = in1 ? tc_exp(x1 - logZ) : 0.0;
    // node potentials of steps t .. t + PF - 1 ride in registers: a step (~1 000 cycles) is shorter than a DRAM round trip
    double xq0[PF], xq1[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      xq0[i] = xq1[i] = NEG;
      if (1 + i < T) ld2n(1 + i, xq0[i], xq1[i]);
    }
    // publish message 0 into buffer 1
    vec[1][w][0][lane] = e0;
    vec[1][w][1][lane] = e1;
    {
      const double ps = quad_sum(e0 + e1);
      const float pm = pmaxf2(xq0[0], xq1[0]);
      if (q == 0) { psum[1][w][g] = ps; pmax[1][w][g] = pm; }
    }
    double cprod = 1.0;
    for (int t = 1; t < T; ++t) {
      const int buf = t & 1;
      __syncthreads();
      double xr0 = NEG, xr1 = NEG;
      if (t + PF < T) ld2n(t + PF, xr0, xr1);
      const double xn0 = xq0[0], xn1 = xq1[0];
      const double c = sum4(buf);
      const double m = max4(buf);
      double d0, d1;
      tile_mv(buf, F, d0, d1);
      const double invc = tc_rcp(c);
      if (!p.fwd_only) st2(wsA, t - 1, e0 * invc, e1 * invc);             // ahat_{t-1}
      e0 = in0 ? d0 * invc * tc_exp(xn0 - m) : 0.0;
      e1 = in1 ? d1 * invc * tc_exp(xn1 - m) : 0.0;
      logZ += m;
      cprod *= c;
      if ((t & 7) == 7) {
        logZ += log(cprod);
        cprod = 1.0;
      }
#pragma unroll
      for (int i = 0; i + 1 < PF; ++i) { xq0[i] = xq0[i + 1]; xq1[i] = xq1[i + 1]; }
      xq0[PF - 1] = xr0;
      xq1[PF - 1] = xr1;
      vec[buf ^ 1][w][0][lane] = e0;
      vec[buf ^ 1][w][1][lane] = e1;
      const double ps = quad_sum(e0 + e1);
      const float pm = pmaxf2(xq0[0], xq1[0]);
      if (q == 0) { psum[buf ^ 1][w][g] = ps; pmax[buf ^ 1][w][g] = pm; }
    }
    __syncthreads();
    const double c = sum4(T & 1);
    cprod *= c;
    logZ += log(cprod) + (double)(T - 1) * gmax;
    if (w == 0 && q == 0 && valid && p.logZ) p.logZ[b] = logZ;
    if (p.fwd_only) return;
    const double invc = 1.0 / c;
    e0 *= invc;
    e1 *= invc;                                                            // ahat_{T-1} = E[z_{T-1}]
    st2(Ez, T - 1, e0, e1);
    if (T == 1 && p.Ez0) st2(p.Ez0 + (size_t)b * K + s0, 0, e0, e1);
  }
  if (T == 1) return;

  // ------------------------------------------------ backward -----------------------------------------------
  double F[8];              // F[s] = P'[k][j] = P[8 w + g][8 (s>>1) + 2 q + (s&1)]
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int k = 8 * (s >> 1) + 2 * q + (s & 1), j = 8 * w + g;
    F[s] = (k < K && j < K) ? exp(p.log_A[j * K + k] - gmax) : 0.0;
  }
  // prologue: U_{T-1} = exp(node_{T-1} - max), s_{T-1} := 1; the maximum needs its own exchange
  double xl0, xl1;
  ld2n(T - 1, xl0, xl1);
  __syncthreads();                                        // everybody is done with the forward buffers
  {
    const float pm = pmaxf2(xl0, xl1);
    if (q == 0) pmax[0][w][g] = pm;
  }
  __syncthreads();
  const double mm = max4(0);
  __syncthreads();                                        // pmax[0] is rewritten below when (T-2) is even
  const int b0 = (T - 2) & 1;
  vec[b0][w][0][lane] = in0 ? tc_exp(xl0 - mm) : 0.0;
  vec[b0][w][1][lane] = in1 ? tc_exp(xl1 - mm) : 0.0;
  // node[t - i], ahat_{t-i} (this lane stored them) for i < PB ride in registers
  auto lda = [&](int t, double& v0, double& v1) {
    v0 = (in0 && valid) ? wsA[(size_t)t * K] : (in0 ? 1.0 / K : 0.0);
    v1 = (in1 && valid) ? wsA[(size_t)t * K + 1] : (in1 ? 1.0 / K : 0.0);
  };
  double qn0[PB], qn1[PB], qa0[PB], qa1[PB];
#pragma unroll
  for (int i = 0; i < PB; ++i) {
    qn0[i] = qn1[i] = NEG;
    qa0[i] = qa1[i] = 0.0;
    if (T - 2 - i >= 0) {
      ld2n(T - 2 - i, qn0[i], qn1[i]);
      lda(T - 2 - i, qa0[i], qa1[i]);
    }
  }
  {
    const float pm = pmaxf2(qn0[0], qn1[0]);
    if (q == 0) { psum[b0][w][g] = 0.25; pmax[b0][w][g] = pm; }
  }
  double prd0 = 0.0, prd1 = 0.0, uw0 = 0.0, uw1 = 0.0;   // finished one step late
  for (int t = T - 2; t >= 0; --t) {
    const int buf = t & 1;
    __syncthreads();
    double xp0 = NEG, xp1 = NEG, ap0 = 0.0, ap1 = 0.0;   // prefetch for iteration t - PB
    if (t - PB >= 0) {
      ld2n(t - PB, xp0, xp1);
      lda(t - PB, ap0, ap1);
    }
    const double xn0 = qn0[0], xn1 = qn1[0], a0 = qa0[0], a1 = qa1[0];
    const double inv = tc_rcp(sum4(buf));                 // 1 / s_{t+1}
    const double m = max4(buf);                           // max of node_t
    double d0, d1;
    tile_mv(buf, F, d0, d1);                              // (P U_{t+1}), own tile
    if (t < T - 2) {                                      // finish step t + 1
      st2(Ez, t + 1, prd0 * inv, prd1 * inv);
      st2(wsW, t + 1, uw0 * inv, uw1 * inv);
    }
    const double B0 = d0 * inv, B1 = d1 * inv;            // bu_t
    prd0 = a0 * B0;
    prd1 = a1 * B1;
    uw0 = vec[buf][w][0][lane] * inv;                     // w_{t+1} = U_{t+1} / s_{t+1}
    uw1 = vec[buf][w][1][lane] * inv;
    vec[buf ^ 1][w][0][lane] = in0 ? tc_exp(xn0 - m) * B0 : 0.0;             // U_t
    vec[buf ^ 1][w][1][lane] = in1 ? tc_exp(xn1 - m) * B1 : 0.0;
#pragma unroll
    for (int i = 0; i + 1 < PB; ++i) { qn0[i] = qn0[i + 1]; qn1[i] = qn1[i + 1]; qa0[i] = qa0[i + 1]; qa1[i] = qa1[i + 1]; }
    qn0[PB - 1] = xp0; qn1[PB - 1] = xp1; qa0[PB - 1] = ap0; qa1[PB - 1] = ap1;
    {
      const double ps = quad_sum(prd0 + prd1);
      const float pm = pmaxf2(qn0[0], qn1[0]);            // node[t-1]: its maximum travels with U_t
      if (q == 0) { psum[buf ^ 1][w][g] = ps; pmax[buf ^ 1][w][g] = pm; }
    }
  }
  __syncthreads();
  {
    const double inv0 = 1.0 / sum4(1);                    // s_0 (published into buffer 0 ^ 1)
    st2(Ez, 0, prd0 * inv0, prd1 * inv0);
    st2(wsW, 0, uw0 * inv0, uw1 * inv0);
    if (p.Ez0) st2(p.Ez0 + (size_t)b * K + s0, 0, prd0 * inv0, prd1 * inv0);
  }
}

// ---- forward-backward sweep, EIGHT warps per 8 sequences ----------------------------------------------------------
// Same exchange scheme as the four-warp sweep; warps 2w and 2w+1 both form n-tile w (the fp64 tensor pipe is 12 % busy:
// the duplicate mma are free) and each keeps ONE of the two accumulator elements, so a lane carries one state: half the
// scalar work per warp-step (one exp, one store stream) and twice the warps per scheduler at a given batch.
template <int PB>
__global__ void __launch_bounds__(256) hmm_tc_sweep8_kernel(const HmmTcParams p) {
  constexpr int PF = SVAE_HMM_TC_PF;
  __shared__ __align__(16) double vec[2][4][2][32];      // [buffer][n-tile][element][lane]
  __shared__ double psum[2][8][8];                       // [buffer][warp][row]: partial row sums
  __shared__ float pmax[2][8][8];                        // partial row maxima of the next node potentials
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, w = wi >> 1, el = wi & 1, g = lane >> 2, q = lane & 3;
  const int K = p.K, T = p.T;
  const long long seq = (long long)blockIdx.x * 8 + g;
  const long long b = seq < p.B ? seq : (long long)p.B - 1;
  const bool valid = seq < p.B && (!p.active || p.active[b] != 0);
  const double NEG = -INFINITY;
  const double gmax = tc_gmax(p.log_A, K, lane);
  const int s0 = 8 * w + 2 * q + el;                      // this lane's state
  const bool in0 = s0 < K;

  const double* node = p.node + (size_t)b * T * K + s0;
  double* Ez = p.Ez ? p.Ez + (size_t)b * T * K + s0 : nullptr;
  double* wsA = p.wsA ? p.wsA + (size_t)b * T * K + s0 : nullptr;
  double* wsW = p.wsW ? p.wsW + (size_t)b * T * K + s0 : nullptr;

  auto ldn = [&](int t) { return in0 ? __ldg(node + (size_t)t * K) : NEG; };
  auto st1 = [&](double* base, int t, double v) {
    if (valid && in0) base[(size_t)t * K] = v;
  };
  auto pmaxf1 = [&](double x) { return quad_maxf((float)x); };
  auto tile_mv = [&](int buf, const double (&F)[8]) {     // element `el` of the own n-tile
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
    for (int s = 0; s < 8; s += 2) {
      hmm_dmma(c0, c1, vec[buf][s >> 1][0][lane], F[s]);
      hmm_dmma(c2, c3, vec[buf][s >> 1][1][lane], F[s + 1]);
    }
    return el ? c1 + c3 : c0 + c2;
  };
  auto sum8 = [&](int buf) {
    return ((psum[buf][0][g] + psum[buf][1][g]) + (psum[buf][2][g] + psum[buf][3][g])) +
           ((psum[buf][4][g] + psum[buf][5][g]) + (psum[buf][6][g] + psum[buf][7][g]));
  };
  auto max8 = [&](int buf) {
    return (double)fmaxf(fmaxf(fmaxf(pmax[buf][0][g], pmax[buf][1][g]), fmaxf(pmax[buf][2][g], pmax[buf][3][g])),
                         fmaxf(fmaxf(pmax[buf][4][g], pmax[buf][5][g]), fmaxf(pmax[buf][6][g], pmax[buf][7][g])));
  };

  // ------------------------------------------------ forward ------------------------------------------------
  double e0;                // this lane's entry of the (unnormalised) message
  {
    double F[8];            // F[s] = P[8 (s>>1) + 2 q + (s&1)][8 w + g]
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int i = 8 * (s >> 1) + 2 * q + (s & 1), j = 8 * w + g;
      F[s] = (i < K && j < K) ? exp(p.log_A[i * K + j] - gmax) : 0.0;
    }
    double x0 = ldn(0);
    x0 = in0 ? p.log_pi[s0] + x0 : NEG;
    {
      const float pm = pmaxf1(x0);
      if (q == 0) pmax[0][wi][g] = pm;
    }
    __syncthreads();
    double logZ = max8(0);
    e0 = in0 ? tc_exp(x0 - logZ) : 0.0;
    double xq0[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) xq0[i] = (1 + i < T) ? ldn(1 + i) : NEG;
    vec[1][w][el][lane] = e0;
    {
      const double ps = quad_sum(e0);
      const float pm = pmaxf1(xq0[0]);
      if (q == 0) { psum[1][wi][g] = ps; pmax[1][wi][g] = pm; }
    }
    double cprod = 1.0;
    for (int t = 1; t < T; ++t) {
      const int buf = t & 1;
      __syncthreads();
      const double xr0 = (t + PF < T) ? ldn(t + PF) : NEG;
      const double xn0 = xq0[0];
      const double c = sum8(buf);
      const double m = max8(buf);
      const double d0 = tile_mv(buf, F);
      const double invc = tc_rcp(c);
      if (!p.fwd_only) st1(wsA, t - 1, e0 * invc);                        // ahat_{t-1}
      e0 = in0 ? d0 * invc * tc_exp(xn0 - m) : 0.0;
      logZ += m;
      cprod *= c;
      if ((t & 7) == 7) {
        logZ += log(cprod);
        cprod = 1.0;
      }
#pragma unroll
      for (int i = 0; i + 1 < PF; ++i) xq0[i] = xq0[i + 1];
      xq0[PF - 1] = xr0;
      vec[buf ^ 1][w][el][lane] = e0;
      const double ps = quad_sum(e0);
      const float pm = pmaxf1(xq0[0]);
      if (q == 0) { psum[buf ^ 1][wi][g] = ps; pmax[buf ^ 1][wi][g] = pm; }
    }
    __syncthreads();
    const double c = sum8(T & 1);
    cprod *= c;
    logZ += log(cprod) + (double)(T - 1) * gmax;
    if (wi == 0 && q == 0 && valid && p.logZ) p.logZ[b] = logZ;
    if (p.fwd_only) return;
    e0 *= 1.0 / c;                                                         // ahat_{T-1} = E[z_{T-1}]
    st1(Ez, T - 1, e0);
    if (T == 1 && p.Ez0) st1(p.Ez0 + (size_t)b * K + s0, 0, e0);
  }
  if (T == 1) return;

  // ------------------------------------------------ backward -----------------------------------------------
  double F[8];              // F[s] = P'[k][j] = P[8 w + g][8 (s>>1) + 2 q + (s&1)]
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int k = 8 * (s >> 1) + 2 * q + (s & 1), j = 8 * w + g;
    F[s] = (k < K && j < K) ? exp(p.log_A[j * K + k] - gmax) : 0.0;
  }
  const double xl0 = ldn(T - 1);
  __syncthreads();                                        // everybody is done with the forward buffers
  {
    const float pm = pmaxf1(xl0);
    if (q == 0) pmax[0][wi][g] = pm;
  }
  __syncthreads();
  const double mm = max8(0);
  __syncthreads();                                        // pmax[0] is rewritten below when (T-2) is even
  const int b0 = (T - 2) & 1;
  vec[b0][w][el][lane] = in0 ? tc_exp(xl0 - mm) : 0.0;
  auto lda = [&](int t) { return (in0 && valid) ? wsA[(size_t)t * K] : (in0 ? 1.0 / K : 0.0); };
  double qn0[PB], qa0[PB];
#pragma unroll
  for (int i = 0; i < PB; ++i) {
    qn0[i] = NEG;
    qa0[i] = 0.0;
    if (T - 2 - i >= 0) {
      qn0[i] = ldn(T - 2 - i);
      qa0[i] = lda(T - 2 - i);
    }
  }
  {
    const float pm = pmaxf1(qn0[0]);
    if (q == 0) { psum[b0][wi][g] = 0.125; pmax[b0][wi][g] = pm; }
  }
  double prd0 = 0.0, uw0 = 0.0;                           // finished one step late
  for (int t = T - 2; t >= 0; --t) {
    const int buf = t & 1;
    __syncthreads();
    double xp0 = NEG, ap0 = 0.0;                          // prefetch for iteration t - PB
    if (t - PB >= 0) {
      xp0 = ldn(t - PB);
      ap0 = lda(t - PB);
    }
    const double xn0 = qn0[0], a0 = qa0[0];
    const double inv = tc_rcp(sum8(buf));                 // 1 / s_{t+1}
    const double m = max8(buf);                           // max of node_t
    const double d0 = tile_mv(buf, F);                    // (P U_{t+1}), own element
    if (t < T - 2) {                                      // finish step t + 1
      st1(Ez, t + 1, prd0 * inv);
      st1(wsW, t + 1, uw0 * inv);
    }
    const double B0 = d0 * inv;                           // bu_t
    prd0 = a0 * B0;
    uw0 = vec[buf][w][el][lane] * inv;                    // w_{t+1} = U_{t+1} / s_{t+1}
    vec[buf ^ 1][w][el][lane] = in0 ? tc_exp(xn0 - m) * B0 : 0.0;         // U_t
#pragma unroll
    for (int i = 0; i + 1 < PB; ++i) { qn0[i] = qn0[i + 1]; qa0[i] = qa0[i + 1]; }
    qn0[PB - 1] = xp0;
    qa0[PB - 1] = ap0;
    {
      const double ps = quad_sum(prd0);
      const float pm = pmaxf1(qn0[0]);                    // node[t-1]: its maximum travels with U_t
      if (q == 0) { psum[buf ^ 1][wi][g] = ps; pmax[buf ^ 1][wi][g] = pm; }
    }
  }
  __syncthreads();
  {
    const double inv0 = 1.0 / sum8(1);                    // s_0 (published into buffer 0 ^ 1)
    st1(Ez, 0, prd0 * inv0);
    st1(wsW, 0, uw0 * inv0);
    if (p.Ez0) st1(p.Ez0 + (size_t)b * K + s0, 0, prd0 * inv0);
  }
}

// ---- transition counts: one warp per sequence, contraction over time on the tensor cores -------------------------
//   Etrans[i][j] = P[i][j] * sum_{t < T-1} ahat_t[i] (w_{t+1}/s_t)[j]
#ifndef SVAE_HMM_TC_TRANS_WARPS
#define SVAE_HMM_TC_TRANS_WARPS 4
#endif
constexpr int kTransWarps = SVAE_HMM_TC_TRANS_WARPS;
__global__ void __launch_bounds__(32 * kTransWarps) hmm_tc_trans_kernel(int B, int T, int K, const double* log_A,
                                                                        const int* active, const double* wsA,
                                                                        const double* wsW, double* Etrans) {
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const long long b = (long long)blockIdx.x * kTransWarps + (threadIdx.x >> 5);
  if (b >= B) return;                                   // warp-uniform
  if (active && active[b] == 0) return;
  const double gmax = tc_gmax(log_A, K, lane);
  const double* A = wsA + (size_t)b * T * K;
  const double* W = wsW + (size_t)b * T * K;
  double acc[4][4][2];
#pragma unroll
  for (int it = 0; it < 4; ++it)
#pragma unroll
    for (int jt = 0; jt < 4; ++jt) acc[it][jt][0] = acc[it][jt][1] = 0.0;
  bool ing[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) ing[it] = (8 * it + g) < K;
  const int steps = T - 1;                              // rows 0 .. T-2 of both arrays
  double a[4], w[4];
  auto fetch = [&](int t0, double (&av)[4], double (&wv)[4]) {
    const int t = t0 + q;
    const bool ok = t < steps;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const bool ld = ok && ing[it];
      av[it] = ld ? A[(size_t)t * K + 8 * it + g] : 0.0;        // A[g][q] = ahat_{t0+q}[8 it + g]
      wv[it] = ld ? W[(size_t)t * K + 8 * it + g] : 0.0;        // B[q][g] = (w/s)_{t0+q}[8 jt + g]
    }
  };
#if SVAE_HMM_TC_TRANS_AHEAD == 2
  double a1[4], w1[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) a[it] = w[it] = a1[it] = w1[it] = 0.0;
  if (steps > 0) fetch(0, a, w);
  if (steps > 4) fetch(4, a1, w1);
  for (int t0 = 0; t0 < steps; t0 += 4) {
    double an[4], wn[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) an[it] = wn[it] = 0.0;
    if (t0 + 8 < steps) fetch(t0 + 8, an, wn);
#pragma unroll
    for (int it = 0; it < 4; ++it)
#pragma unroll
      for (int jt = 0; jt < 4; ++jt) hmm_dmma(acc[it][jt][0], acc[it][jt][1], a[it], w[jt]);
#pragma unroll
    for (int it = 0; it < 4; ++it) { a[it] = a1[it]; w[it] = w1[it]; a1[it] = an[it]; w1[it] = wn[it]; }
  }
#else
#pragma unroll
  for (int it = 0; it < 4; ++it) a[it] = w[it] = 0.0;
  if (steps > 0) fetch(0, a, w);
  for (int t0 = 0; t0 < steps; t0 += 4) {
    double an[4], wn[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) an[it] = wn[it] = 0.0;
    if (t0 + 4 < steps) fetch(t0 + 4, an, wn);
#pragma unroll
    for (int it = 0; it < 4; ++it)
#pragma unroll
      for (int jt = 0; jt < 4; ++jt) hmm_dmma(acc[it][jt][0], acc[it][jt][1], a[it], w[jt]);
#pragma unroll
    for (int it = 0; it < 4; ++it) { a[it] = an[it]; w[it] = wn[it]; }
  }
#endif
  double* out = Etrans + (size_t)b * K * K;
#pragma unroll
  for (int it = 0; it < 4; ++it)
#pragma unroll
    for (int jt = 0; jt < 4; ++jt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 8 * it + g, j = 8 * jt + 2 * q + e;
        if (i < K && j < K) out[i * K + j] = exp(log_A[i * K + j] - gmax) * acc[it][jt][e];
      }
}

}  // namespace svae

// workspace doubles of the tensor-core path
size_t svae_hmm_tc_workspace_bytes(int B, int T, int K) { return sizeof(double) * 2 * (size_t)B * (size_t)T * (size_t)K; }

bool svae_hmm_tc_enabled() {
  static const bool on = [] { const char* e = getenv("SVAE_HMM_TC"); return !(e && atoi(e) == 0); }();
  return on;
}

// called by svae_hmm_estep_wide (hmm_wide.cu) for SVAE_HMM_TC_MIN_K <= K <= 32 when the caller brought the workspace
int svae_hmm_estep_tc_launch(int B, int T, int K, bool fwd_only, const double* log_pi, const double* log_A, const double* node,
                             const int* active, double* logZ, double* Ez0, double* Etrans, double* Ez, void* workspace,
                             cudaStream_t s) {
  svae::HmmTcParams p;
  p.B = B; p.T = T; p.K = K; p.fwd_only = fwd_only ? 1 : 0;
  p.log_pi = log_pi; p.log_A = log_A; p.node = node; p.active = active;
  p.logZ = logZ; p.Ez0 = Ez0; p.Ez = Ez;
  p.wsA = static_cast<double*>(workspace);              // NULL is fine for the log-normaliser alone
  p.wsW = workspace ? p.wsA + (size_t)B * T * K : nullptr;
  // four warps per 8 sequences (one n-tile each) unless SVAE_HMM_TC_WARPS=1 asks for the one-warp sweep
  static const int warps = [] { const char* e = getenv("SVAE_HMM_TC_WARPS"); return e ? atoi(e) : 4; }();
  if (warps == 1)
    svae::hmm_tc_sweep_kernel<<<(unsigned)((B + 7) / 8), 32, 0, s>>>(p);
  else if (warps == 8)
    svae::hmm_tc_sweep8_kernel<2><<<(unsigned)((B + 7) / 8), 256, 0, s>>>(p);
  else if (warps == 81)
    svae::hmm_tc_sweep8_kernel<1><<<(unsigned)((B + 7) / 8), 256, 0, s>>>(p);
  else {
    // backward prefetch depth: 2 steps while every CTA of the grid is resident whatever its register count (measured
    // 0.965 -> 0.930 ms at 2048 sequences), 1 beyond (the extra registers cost occupancy: 4.41 vs 4.54 ms at 16 384)
    const unsigned grid = (unsigned)((B + 7) / 8);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (grid <= 4u * (unsigned)sms)
      svae::hmm_tc_sweep4_kernel<2><<<grid, 128, 0, s>>>(p);
    else
      svae::hmm_tc_sweep4_kernel<1><<<grid, 128, 0, s>>>(p);
  }
  g_svae_launches.fetch_add(1);
  if (!fwd_only && Etrans) {
    svae::hmm_tc_trans_kernel<<<(unsigned)((B + svae::kTransWarps - 1) / svae::kTransWarps), 32 * svae::kTransWarps, 0, s>>>(
        B, T, K, log_A, active, p.wsA, p.wsW, Etrans);
    g_svae_launches.fetch_add(1);
  }
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}
