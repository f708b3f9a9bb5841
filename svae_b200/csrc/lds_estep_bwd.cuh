// lds_estep_bwd.cuh -- reverse-mode (VJP) of the fused LDS E-step w.r.t. the node potentials, sm_100a.
//
// What it replaces (reference, one sequence per call):
//   svae/lds/cython_lds_inference.pyx:357-409  natural_sample_backward_grad
//   svae/lds/cython_lds_inference.pyx:236-306  natural_smoother_general_grad (+ _compute_stats_grad :213-234)
//   svae/lds/cython_lds_inference.pyx:92-145   natural_filter_grad
//   and the per-step VJPs in svae/lds/cython_gaussian_grads.pxd:84-120,146-162,190-204,222-238,298-426,
//   456-487,510-528 with svae/cython_linalg_grads.pxd:10-118, composed as lds_inference.py:26-39 wires them.
//
// B200 design: the adjoint of THIS library's forward formulation (lds_estep.cuh), not a translation of the
// reference's VJPs (the gradient is unique; parity is checked against the reference's VJP outputs).
// With G_t, W_t = M_t^-1, m_t, K_t = L_t^-1 and the smoothed Sigma_t saved by the forward kernel:
//
//   sweep 1 (forward in time) -- adjoint of the smoother/sampler recursion
//       Sigma_t = W_t + C_t G_t',  C_t = G_t Sigma_{t+1},  v_t = base_t + G_t v_{t+1}  (v = mu, x^s)
//     carries (Sb = dL/dSigma_t, vb = dL/dv_t), and reduces everything that depends on step t alone to
//       Q_t  = -W (Sb W + Gb G') - (W mb) m' - g/2 (W + m m') + sum_s K' Phi(-eps_s u_s') K,  u_s = K xb_s
//       f_t  = W mb + g m                         (mb = mub + sum_s xb_s,  g = dL/dlognorm)
//     with Cb = B + Sb G,  Gb = Sb' C + Cb Sigma_{t+1}' + sum_v vb v_{t+1}',  Sb_{t+1} = G' Cb,  vb_{t+1} = G' vb;
//   sweep 2 (backward in time) -- adjoint of the filter recursion, a smoother-shaped recursion itself
//       Pb_t = Q_t + G_t Pb_{t+1} G_t' - (G_t hb_{t+1}) m_t',   hb_t = f_t + G_t hb_{t+1}
//     and the answers are  dL/dJ_node[t] = -2 diag(Pb_t),  dL/dh_node[t] = hb_t.
// (Derivation + NumPy prototype: scripts/dev_adjoint_numpy.py.)  Lanes = matrix columns as in the forward.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "async_copy.cuh"
#include "lds_estep.cuh"

namespace svae {

struct LdsBwdParams {
  int B, T, S, flags;
  // forward results this pass re-reads
  const double *E_node_x, *samples, *eps;
  // cotangents (any may be NULL = zero)
  const double *g_samples;                                 // (B,T,S,n)
  const double *g_E_init_xx, *g_E_init_x;                  // (B,n,n), (B,n)
  const double *g_E_pair_xx, *g_E_pair_xxn, *g_E_pair_xnxn;  // (B,n,n) homogeneous sums
  const double *g_Exx_t, *g_Exxn_t;                        // (B,T,n,n), (B,T-1,n,n) per-time moments
  const double *g_E_node_xx, *g_E_node_x;                  // (B,T,n)
  const double *g_lognorm;                                 // (B)
  // outputs
  double *g_node_J, *g_node_h;                             // (B,T,n)
  double *g_node_J_dense;                                  // optional (B,T,n,n): raw Pbar_t, caller symmetrises
  // message ports (the reference's three primitives differentiated one by one, lds_inference.py:26-39)
  double *g_msg_M, *g_msg_h;                               // out (sweep 1): cotangent of M_t = filtered precision + J11
                                                           //   (information form, raw) (B,T,n,n) and of h_filtered (B,T,n)
  const double *in_g_M_pred, *in_g_h_pred;                 // in (sweep 2 only): cotangents of the PREDICTION messages,
  const double *in_g_M_filt, *in_g_h_filt;                 //   of the FILTERED messages (information form) (B,T,n,n), (B,T,n)
  double* stash;                                           // forward workspace (SAVE layout); overwritten
};
#define SVAE_LDS_BWD_SWEEP1_ONLY 8     // adjoint of the smoother / sampler alone: message cotangents out, no filter adjoint
#define SVAE_LDS_BWD_SWEEP2_ONLY 16    // adjoint of the filter alone: message cotangents in

// shared-memory doubles per sequence tile of the VJP kernel
__host__ __device__ inline int lds_bwd_tile_doubles(int N, int S) {
  const int NP = lds_np(N);
  const int REC = (3 * N + 1 + S) * NP;   // G | W | m | b | K   (Sigma_{t+1} is fetched separately)
  return 2 /*mbarriers*/ + 2 * (REC + N * NP) /*ring: record + Sigma_{t+1}*/ + 4 * N * NP /*XA XB X0 Bm*/ +
         3 * N * NP /*SbS CbS GbS*/ + (4 + 1 + (1 + S) + 2) * NP /*mean window(4), vtmp, vbS, wm, us*/ +
         2 * (2 + 3 * S) * NP /*input ring: g_node_x | g_node_xx | g_samples | x_{t+1} | eps*/;
}

// PER_T: cotangents arrive for the per-time second moments (inhomogeneous / SLDS use); the homogeneous
// instantiation carries none of that code.
template <int N, int LPS, bool PER_T>
__global__ void __launch_bounds__(32, 8) lds_estep_bwd_kernel(const LdsBwdParams p) {
  constexpr int SPW = 32 / LPS;
  constexpr int NP = lds_np(N);
  constexpr int NN = N * N;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPS, c = lane % LPS;
  const int S = p.S, T = p.T;
  const long long seq = ((long long)blockIdx.x * nwarps + warp) * SPW + sub;
  const bool valid = seq < p.B;
  const long long b = valid ? seq : (long long)p.B - 1;
  const bool mat = c < N, vec = c == N, smp = (c > N) && (c <= N + S);
  const int s = smp ? c - N - 1 : 0;
  const int RS = lds_record_stride(N, S, true);
  const int REC = (3 * N + 1 + S) * NP;
  const int SLOT = REC + N * NP;

  double* tile = smem + (size_t)(warp * SPW + sub) * lds_bwd_tile_doubles(N, S);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tile);
  double* ring = tile + 2;               // 2 x { record prefix | Sigma_{t+1} }
  double* XA = ring + 2 * SLOT;          // sym(g_E_pair_xx)
  double* XB = XA + N * NP;              // sym(g_E_pair_xnxn)
  double* X0 = XB + N * NP;              // sym(g_E_init_xx)
  double* Bm = X0 + N * NP;              // g_E_pair_xxn (not symmetric)
  double* SbS = Bm + N * NP;             // SbS[k*NP+i] = Sb[i][k]
  double* CbS = SbS + N * NP;            // CbS[k*NP+i] = Cb[i][k]
  double* GbS = CbS + N * NP;            // GbS[k*NP+i] = Gb[i][k]   (sweep 2: Pb)
  double* muw = GbS + N * NP;            // mean window: slot (t mod 4) holds mu_t
  double* vtmp = muw + 4 * NP;           // per-lane scalars exchanged through smem
  double* vbS = vtmp + NP;               // vbS[v*NP+i]: vb of vector column v (0 = mu, 1+s = x^s)
  double* wm = vbS + (1 + S) * NP;       // W mb
  double* us = wm + NP;                  // u = K xb_s
  const int IN = (2 + 3 * S) * NP;       // per-step inputs staged one step ahead by cp.async:
  double* iring = us + NP;               //   2 x { g_node_x | g_node_xx | g_samples[S] | x_{t+1}[S] | eps_t[S] }

  if (c == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  constexpr bool per_t = PER_T;
  const bool s2only = (p.flags & SVAE_LDS_BWD_SWEEP2_ONLY) != 0;
  double* stash = p.stash + (size_t)b * T * RS;
  const double* Ex = p.E_node_x + (size_t)b * T * N;
  const double g_ln = p.g_lognorm ? p.g_lognorm[b] : 0.0;
  if (!s2only) {
  // stage the (symmetrised) cotangents of the global statistics
  if (mat) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const size_t o = (size_t)b * NN, ic = i * N + c, ci = c * N + i;
      XA[i * NP + c] = p.g_E_pair_xx ? 0.5 * (p.g_E_pair_xx[o + ic] + p.g_E_pair_xx[o + ci]) : 0.0;
      XB[i * NP + c] = p.g_E_pair_xnxn ? 0.5 * (p.g_E_pair_xnxn[o + ic] + p.g_E_pair_xnxn[o + ci]) : 0.0;
      X0[i * NP + c] = p.g_E_init_xx ? 0.5 * (p.g_E_init_xx[o + ic] + p.g_E_init_xx[o + ci]) : 0.0;
      Bm[i * NP + c] = p.g_E_pair_xxn ? p.g_E_pair_xxn[o + ic] : 0.0;
    }
  }
  __syncthreads();

  // forward records were written through the generic proxy by another kernel: visible to TMA already.
  auto fetch1 = [&](int t) {   // sweep 1, time t -> slot t&1: record t prefix + Sigma_{t+1}
#if !SVAE_LDS_TMA   // plain-load variant (sanitizer runs: racecheck cannot order async-proxy writes against generic reads)
    {
      double* dst = ring + (t & 1) * SLOT;
      const double* src = stash + (size_t)t * RS;
      for (int e = c; e < REC; e += LPS) dst[e] = src[e];
      if (t + 1 < T) {
        const double* s2 = stash + (size_t)(t + 1) * RS + (3 * N + 1 + S) * NP;
        for (int e = c; e < N * NP; e += LPS) dst[REC + e] = s2[e];
      }
      return;
    }
#endif
    if (c == 0) {
      const uint32_t nb1 = (uint32_t)(REC * sizeof(double)), nb2 = (uint32_t)(N * NP * sizeof(double));
      const bool has_next = t + 1 < T;
      mbar_arrive_expect_tx(&bars[t & 1], nb1 + (has_next ? nb2 : 0u));
      tma_load_1d(ring + (t & 1) * SLOT, stash + (size_t)t * RS, nb1, &bars[t & 1]);
      if (has_next)
        tma_load_1d(ring + (t & 1) * SLOT + REC, stash + (size_t)(t + 1) * RS + (3 * N + 1 + S) * NP, nb2, &bars[t & 1]);
    }
  };
  fetch1(0);
  if (T > 1) fetch1(1);

  // per-step inputs (cotangents, next sample, noise) and the mean window arrive by cp.async one step ahead
  for (int e = c; e < 2 * IN; e += LPS) iring[e] = 0.0;
  __syncwarp();
  auto stage_in = [&](int t) {    // inputs of step t -> input-ring slot t&1;  mu_{t+1} -> window slot (t+1)%4
    if (mat && t < T) {
      double* dst = iring + (t & 1) * IN;
      const size_t bt = (size_t)b * T + t;
      if (p.g_E_node_x) cp_async8(dst + c, p.g_E_node_x + bt * N + c);
      if (p.g_E_node_xx) cp_async8(dst + NP + c, p.g_E_node_xx + bt * N + c);
      for (int ss = 0; ss < S; ++ss) {
        if (p.g_samples) cp_async8(dst + (2 + ss) * NP + c, p.g_samples + (bt * S + ss) * N + c);
        if (t + 1 < T) cp_async8(dst + (2 + S + ss) * NP + c, p.samples + ((bt + 1) * S + ss) * N + c);
        cp_async8(dst + (2 + 2 * S + ss) * NP + c, p.eps + (bt * S + ss) * N + c);
      }
      if (t + 1 < T) cp_async8(muw + ((t + 1) % 4) * NP + c, Ex + (size_t)(t + 1) * N + c);
    }
    cp_async_commit();
  };
  if (mat) cp_async8(muw + c, Ex + c);
  stage_in(0);

  double Sb[N];    // mat: column c of Sb;  vec/smp: the vector cotangent vb
#pragma unroll
  for (int i = 0; i < N; ++i) Sb[i] = 0.0;

  // ================================ sweep 1: forward in time ==================================
  for (int t = 0; t < T; ++t) {
    const bool first = (t == 0), last = (t == T - 1);
    const double* rec = ring + (t & 1) * SLOT;
    const double* G = rec;
    const double* W = rec + N * NP;
    const double* mrow = rec + 2 * N * NP;
    const double* K = rec + (2 * N + 1 + S) * NP;
    const double* SigN = rec + REC;
    const double* mu_t = muw + (t % 4) * NP;
    const double* mu_n = muw + ((t + 1) % 4) * NP;
    const double* mu_p = muw + ((t + 3) % 4) * NP;
    const double* inp = iring + (t & 1) * IN;     // this step's staged inputs
    const double* xn = inp + (2 + S) * NP;        // x_{t+1}[s][i]
    cp_async_wait<0>();
    __syncwarp();
    stage_in(t + 1);                               // lands while this step computes

    // ---- seeds -----------------------------------------------------------------------------
    if (mat) {
      double seedv = inp[c];
      if (first && p.g_E_init_x) seedv += p.g_E_init_x[(size_t)b * N + c];
      double am = 0.0, bn = 0.0, bp = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double a_ic = (last ? 0.0 : XA[i * NP + c]) + (first ? X0[i * NP + c] : XB[i * NP + c]);
        if (per_t && p.g_Exx_t) {
          const size_t o = ((size_t)b * T + t) * NN;
          a_ic += 0.5 * (p.g_Exx_t[o + i * N + c] + p.g_Exx_t[o + c * N + i]);
        }
        if (i == c) a_ic += inp[NP + c];
        Sb[i] += a_ic;
        am = fma(a_ic, mu_t[i], am);                                   // (A_s mu_t)_c
        double bm_ci = Bm[c * NP + i], bm_ic = Bm[i * NP + c];   // homogeneous sums' cotangent (zeros if absent)
        if (per_t && p.g_Exxn_t) {                                // plus the per-time one
          bm_ci += last ? 0.0 : p.g_Exxn_t[((size_t)b * (T - 1) + t) * NN + c * N + i];
          bm_ic += first ? 0.0 : p.g_Exxn_t[((size_t)b * (T - 1) + t - 1) * NN + i * N + c];
        }
        if (!last) bn = fma(bm_ci, mu_n[i], bn);                       // (B mu_{t+1})_c
        if (!first) bp = fma(bm_ic, mu_p[i], bp);                      // (B' mu_{t-1})_c
      }
      vtmp[c] = seedv + 2.0 * am + bn + bp;
#pragma unroll
      for (int i = 0; i < N; ++i) SbS[c * NP + i] = Sb[i];
    } else if (smp) {
#pragma unroll
      for (int i = 0; i < N; ++i) Sb[i] += inp[(2 + s) * NP + i];
    }
    __syncwarp();
    if (vec) {
#pragma unroll
      for (int i = 0; i < N; ++i) Sb[i] += vtmp[i];
    }
    if (vec || smp) {
      const int v = vec ? 0 : 1 + s;
#pragma unroll
      for (int i = 0; i < N; ++i) vbS[v * NP + i] = Sb[i];
    }
#if SVAE_LDS_TMA
    mbar_wait(&bars[t & 1], (uint32_t)((t >> 1) & 1));
#endif
    __syncwarp();

    // ---- the per-step reduction ----------------------------------------------------------------
    double inner[N];
#pragma unroll
    for (int i = 0; i < N; ++i) inner[i] = 0.0;
    double Sbn[N];   // next-step carry: G' Cb (mat) / G' vb (vec, smp)
#pragma unroll
    for (int i = 0; i < N; ++i) Sbn[i] = 0.0;
    if (!last) {
      double Cb[N], sig[N], Gb[N];
      if (mat) {
        double C[N];
#pragma unroll
        for (int i = 0; i < N; ++i) sig[i] = SigN[i * NP + c];   // Sigma_{t+1}[:, c] (symmetric)
        // C[:,c] = G Sigma_{t+1}[:,c]   (G is stored transposed: G[k*NP+i] = G_t[i][k])
#pragma unroll
        for (int i = 0; i < N; ++i) C[i] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 g = ld2(&G[k * NP + 2 * q2]);
            C[2 * q2] = fma(g.x, sig[k], C[2 * q2]);
            if (2 * q2 + 1 < N) C[2 * q2 + 1] = fma(g.y, sig[k], C[2 * q2 + 1]);
          }
        }
        // Cb[:,c] = B[:,c] + Sb G[:,c]
#pragma unroll
        for (int i = 0; i < N; ++i)
          Cb[i] = Bm[i * NP + c] + ((per_t && p.g_Exxn_t) ? p.g_Exxn_t[((size_t)b * (T - 1) + t) * NN + i * N + c] : 0.0);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const double gk = G[c * NP + k];   // G_t[k][c]
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 v = ld2(&SbS[k * NP + 2 * q2]);
            Cb[2 * q2] = fma(v.x, gk, Cb[2 * q2]);
            if (2 * q2 + 1 < N) Cb[2 * q2 + 1] = fma(v.y, gk, Cb[2 * q2 + 1]);
          }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) CbS[c * NP + i] = Cb[i];
        // Gb[:,c] = Sb' C[:,c] + sum_v vb_v v_{t+1}[c]   (+ Cb Sigma_{t+1}' once CbS is visible)
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 v = ld2(&SbS[i * NP + 2 * q2]);     // Sb[k][i], k = 2q2, 2q2+1
            acc = fma(v.x, C[2 * q2], acc);
            if (2 * q2 + 1 < N) acc = fma(v.y, C[2 * q2 + 1], acc);
          }
          Gb[i] = acc;
        }
        const double mn_c = mu_n[c];
#pragma unroll
        for (int i = 0; i < N; ++i) Gb[i] = fma(vbS[i], mn_c, Gb[i]);
        for (int ss = 0; ss < S; ++ss) {
          const double xc = xn[ss * NP + c];
#pragma unroll
          for (int i = 0; i < N; ++i) Gb[i] = fma(vbS[(1 + ss) * NP + i], xc, Gb[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) Cb[i] = Sb[i];   // vector lanes carry G' vb
      }
      __syncwarp();   // CbS visible
      if (mat) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 v = ld2(&CbS[k * NP + 2 * q2]);
            Gb[2 * q2] = fma(v.x, sig[k], Gb[2 * q2]);                 // Cb[:,k] Sigma_{t+1}[c][k]
            if (2 * q2 + 1 < N) Gb[2 * q2 + 1] = fma(v.y, sig[k], Gb[2 * q2 + 1]);
          }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) GbS[c * NP + i] = Gb[i];
      }
      // carry: (G' Cb)[i] = sum_k G_t[k][i] Cb[k] = sum_k G[i*NP+k] Cb[k]   (two partial sums per output)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 g = ld2(&G[i * NP + 2 * q2]);
          s0 = fma(g.x, Cb[2 * q2], s0);
          if (2 * q2 + 1 < N) s1 = fma(g.y, Cb[2 * q2 + 1], s1);
        }
        Sbn[i] = s0 + s1;
      }
      __syncwarp();   // GbS visible
      if (mat) {
        // inner[:,c] = (Gb G')[:,c] = sum_k Gb[:,k] G[c][k]
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const double gck = G[k * NP + c];   // G_t[c][k]
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 v = ld2(&GbS[k * NP + 2 * q2]);
            inner[2 * q2] = fma(v.x, gck, inner[2 * q2]);
            if (2 * q2 + 1 < N) inner[2 * q2 + 1] = fma(v.y, gck, inner[2 * q2 + 1]);
          }
        }
      }
    }
    double Q[N];
    if (mat) {
      // inner += (Sb W)[:,c] = sum_k Sb[:,k] W[k][c]
      double mb_dot = 0.0;   // (W mb)_c
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double wk = W[k * NP + c];
        double mbk = vbS[k];
        for (int ss = 0; ss < S; ++ss) mbk += vbS[(1 + ss) * NP + k];
        mb_dot = fma(wk, mbk, mb_dot);
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 v = ld2(&SbS[k * NP + 2 * q2]);
          inner[2 * q2] = fma(v.x, wk, inner[2 * q2]);
          if (2 * q2 + 1 < N) inner[2 * q2 + 1] = fma(v.y, wk, inner[2 * q2 + 1]);
        }
      }
      wm[c] = mb_dot;
      // Q[:,c] = -(W inner)[:,c]
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 w = ld2(&W[i * NP + 2 * q2]);
          acc = fma(w.x, inner[2 * q2], acc);
          if (2 * q2 + 1 < N) acc = fma(w.y, inner[2 * q2 + 1], acc);
        }
        Q[i] = -acc;
      }
    }
    __syncwarp();   // wm visible
    if (mat) {
      const double m_c = mrow[c];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        Q[i] = fma(-wm[i], m_c, Q[i]);                                         // -(W mb) m'
        Q[i] = fma(-0.5 * g_ln, fma(mrow[i], m_c, W[i * NP + c]), Q[i]);       // -g/2 (W + m m')
      }
    }
    // noise adjoint: sum_s K' Phi(-eps_s u_s') K,  u_s = K xb_s
    for (int ss = 0; ss < S; ++ss) {
      if (mat) {
        double u = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) u = fma(K[c * NP + k], vbS[(1 + ss) * NP + k], u);   // K lower: zeros above
        us[c] = u;
      }
      __syncwarp();
      if (mat) {
        const double* ep = inp + (2 + 2 * S + ss) * NP;
        // y_i = -eps_i ( sum_{j<i} u_j K[j][c] + u_i K[i][c] / 2 ),  then  Q[:,c] += K' y
        double run = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double kic = K[i * NP + c];
          const double uk = us[i] * kic;
          const double y = -ep[i] * fma(0.5, uk, run);
          run += uk;
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 kk = ld2(&K[i * NP + 2 * q2]);        // K[i][j], j <= i non-zero
            if (2 * q2 <= i) Q[2 * q2] = fma(kk.x, y, Q[2 * q2]);
            if (2 * q2 + 1 <= i && 2 * q2 + 1 < N) Q[2 * q2 + 1] = fma(kk.y, y, Q[2 * q2 + 1]);
          }
        }
      }
      __syncwarp();
    }
    // ---- write Q_t (over the K block) and f_t (over the first row of W) ------------------------------
    if (mat && valid) {
      double* kt = stash + (size_t)t * RS + (2 * N + 1 + S) * NP;
#pragma unroll
      for (int i = 0; i < N; ++i) kt[i * NP + c] = Q[i];
      stash[(size_t)t * RS + N * NP + c] = fma(g_ln, mrow[c], wm[c]);   // f_t over the (dead) first row of W
      if (p.g_msg_M) {
        const size_t bt = (size_t)b * T + t;
#pragma unroll
        for (int i = 0; i < N; ++i) p.g_msg_M[bt * NN + i * N + c] = Q[i];
        p.g_msg_h[bt * N + c] = fma(g_ln, mrow[c], wm[c]);
      }
    }
    // carry to t+1
#pragma unroll
    for (int i = 0; i < N; ++i) Sb[i] = Sbn[i];
    __syncwarp();
    // ring slot t&1 is free -> fetch t+2
    if (t + 2 < T) fetch1(t + 2);
    __syncwarp();
  }
  }   // !s2only
  else {
    __syncthreads();
  }
  if (p.flags & SVAE_LDS_BWD_SWEEP1_ONLY) return;

  // ================================ sweep 2: backward in time =================================
  // Q_t / f_t were written through the generic proxy: order them before the async-proxy reads
  __threadfence_block();
  fence_proxy_async();
  __syncwarp();
  // barrier phases continue from sweep 1: slot 0 completed ceil(T/2) times, slot 1 floor(T/2) times
  auto fetch2 = [&](int t, int j) {   // j-th step of sweep 2 (t = T-1-j) -> slot j&1
#if !SVAE_LDS_TMA
    {
      double* dst = ring + (j & 1) * SLOT;
      const double* src = stash + (size_t)t * RS;
      for (int e = c; e < REC; e += LPS) dst[e] = src[e];
      return;
    }
#endif
    if (c == 0) {
      const uint32_t nb = (uint32_t)(REC * sizeof(double));
      mbar_arrive_expect_tx(&bars[j & 1], nb);
      tma_load_1d(ring + (j & 1) * SLOT, stash + (size_t)t * RS, nb, &bars[j & 1]);
    }
  };
  fetch2(T - 1, 0);
  if (T > 1) fetch2(T - 2, 1);
  double* PbS = GbS;   // PbS[k*NP+i] = Pb_{t+1}[i][k]
  double* hbS = vtmp;  // hb_{t+1}
  double Pb[N];
  double hb_c = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) Pb[i] = 0.0;
  for (int t = T - 1; t >= 0; --t) {
    const int j = T - 1 - t;
    const double* rec = ring + (j & 1) * SLOT;
    const double* G = rec;
    const double* mrow = rec + 2 * N * NP;
    const double* frow = rec + N * NP;
    const double* Qm = rec + (2 * N + 1 + S) * NP;
    const int used0 = s2only ? 0 : ((j & 1) == 0 ? (T + 1) / 2 : T / 2);
#if SVAE_LDS_TMA
    mbar_wait(&bars[j & 1], (uint32_t)((used0 + (j >> 1)) & 1));
#else
    (void)used0;
    __syncwarp();
#endif
    if (mat) {
      double Pn[N];
      double hn_c;
      const size_t btq = (size_t)b * T + t;
      if (!s2only) {
#pragma unroll
        for (int i = 0; i < N; ++i) Pn[i] = Qm[i * NP + c];
        hn_c = frow[c];
      } else {
        // seeds of the filter adjoint alone: message cotangents + the log-normaliser's own terms -g/2 (W + m m'), g m
        const double* Wm = rec + N * NP;
        const double m_c0 = mrow[c];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double v = -0.5 * g_ln * fma(mrow[i], m_c0, Wm[i * NP + c]);
          if (p.in_g_M_pred) v += p.in_g_M_pred[btq * NN + i * N + c];
          if (p.in_g_M_filt) v += p.in_g_M_filt[btq * NN + i * N + c];
          Pn[i] = v;
        }
        hn_c = g_ln * m_c0 + (p.in_g_h_pred ? p.in_g_h_pred[btq * N + c] : 0.0) + (p.in_g_h_filt ? p.in_g_h_filt[btq * N + c] : 0.0);
      }
      if (t < T - 1) {
        // T1[:,c] = (Pb_{t+1} G')[:,c] = sum_k Pb[:,k] G[c][k];   Ghb = G hb_{t+1}
        double T1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) T1[i] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const double gck = G[k * NP + c];   // G_t[c][k]
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 v = ld2(&PbS[k * NP + 2 * q2]);
            T1[2 * q2] = fma(v.x, gck, T1[2 * q2]);
            if (2 * q2 + 1 < N) T1[2 * q2 + 1] = fma(v.y, gck, T1[2 * q2 + 1]);
          }
        }
        const double m_c = mrow[c];
        // (G T1)[i] = sum_k G_t[i][k] T1[k] and (G hb)[i]: axpy over k on the transposed storage
        double gt1[N], ghb[N];
#pragma unroll
        for (int i = 0; i < N; ++i) gt1[i] = ghb[i] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const double hk = hbS[k];
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 g = ld2(&G[k * NP + 2 * q2]);
            gt1[2 * q2] = fma(g.x, T1[k], gt1[2 * q2]);
            ghb[2 * q2] = fma(g.x, hk, ghb[2 * q2]);
            if (2 * q2 + 1 < N) {
              gt1[2 * q2 + 1] = fma(g.y, T1[k], gt1[2 * q2 + 1]);
              ghb[2 * q2 + 1] = fma(g.y, hk, ghb[2 * q2 + 1]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
          Pn[i] += gt1[i] - ghb[i] * m_c;          // Q + G Pb G' - (G hb) m'
          if (i == c) hn_c += ghb[i];               // f + G hb
        }
      }
#pragma unroll
      for (int i = 0; i < N; ++i) Pb[i] = Pn[i];
      hb_c = hn_c;
      // the node potential enters the FILTERED message only: the prediction message's own cotangent is carried on
      // (it reaches the previous step through the recursion) but is not part of d/d node_t
      const bool sub_pred = s2only && p.in_g_M_pred != nullptr;
      double diag = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i == c) diag = Pb[i] - (sub_pred ? p.in_g_M_pred[btq * NN + i * N + c] : 0.0);
      if (valid) {
        const size_t bt = ((size_t)b * T + t) * N + c;
        if (p.g_node_J) p.g_node_J[bt] = -2.0 * diag;
        if (p.g_node_h) p.g_node_h[bt] = hb_c - ((s2only && p.in_g_h_pred) ? p.in_g_h_pred[bt] : 0.0);
        if (p.g_node_J_dense) {
#pragma unroll
          for (int i = 0; i < N; ++i)
            p.g_node_J_dense[((size_t)b * T + t) * NN + i * N + c] = Pb[i] - (sub_pred ? p.in_g_M_pred[btq * NN + i * N + c] : 0.0);
        }
      }
    }
    __syncwarp();   // everyone is done with PbS / hbS of t+1 and with the ring slot
    if (mat) {
#pragma unroll
      for (int i = 0; i < N; ++i) PbS[c * NP + i] = Pb[i];
      hbS[c] = hb_c;
    }
    if (t >= 2) fetch2(t - 2, j + 2);
    __syncwarp();
  }
}

}  // namespace svae
