// lds_estep_bwd_packed.cuh -- the training-path variant of the reverse-mode kernel (lds_estep_bwd.cuh): same adjoint
// recursions, same workspace record, the lane mapping of lds_estep_packed.cuh.
//
// What it replaces in the reference is what lds_estep_bwd.cuh replaces (cython_lds_inference.pyx:92-145, 236-306,
// 357-409 composed as lds_inference.py:26-39) for the cotangents an SVI step actually produces (svae/svae.py:19-30,
// models/lds.py:35-41): those of the samples, of the node statistics and of the log-normaliser.  The expected GLOBAL
// statistics are consumed un-differentiated by the natural-gradient step (svae.py:31-34 unboxes them), so their
// cotangents -- and with them four n x n matrices per sequence in shared memory plus the mean window -- are absent
// here; calls that do carry them (and the message ports, per-time cotangents, dense node precisions) run the
// general kernel.
//
// Why: the general kernel needs 13.9 KB of shared memory per sequence and 222 registers, so 16 sequences fit on an SM
// and the metric batch takes two waves of a latency-bound kernel (3.8 ms).  Here a sequence owns N lanes (3 per warp at
// n = 10), its vector cotangents are distributed over those lanes, per-lane operands (Sigma_{t+1}[:,c], the per-step
// cotangent elements) come straight from global memory into registers, Cb and Gb share one tile: 7.1 KB per sequence,
// 30 sequences per SM, ONE wave.
#pragma once
#include "lds_estep_bwd.cuh"
#include "lds_estep_packed.cuh"

namespace svae {

__host__ __device__ constexpr int lds_bwd_packed_tile_doubles(int N, int S) {
  const int NP = lds_np(N);
  const int raw = 2 * (3 * N + 1 + S) * NP /*record ring*/ + 2 * N * NP /*SbS, CbS=GbS=PbS*/ + (1 + S) * NP /*vbS*/ +
                  S * NP /*epS*/ + 2 * NP /*wm, us*/;
  return raw + ((NP % 16 - raw % 16) + 16) % 16;
}

#ifndef SVAE_PACKED_BWD_MINB
#define SVAE_PACKED_BWD_MINB(N) ((N) <= 8 ? 12 : (N) <= 10 ? 10 : 8)
#endif

template <int N, int S>
__global__ void __launch_bounds__(32, SVAE_PACKED_BWD_MINB(N)) lds_estep_bwd_packed_kernel(const LdsBwdParams p) {
  constexpr int NP = lds_np(N);
  constexpr int SPW = lds_packed_spw(N);
  constexpr int ACT = SPW * N;
  constexpr unsigned MASK = ACT == 32 ? 0xffffffffu : ((1u << ACT) - 1u);
  constexpr int RS = (4 * N + 1 + S) * NP;          // == lds_record_stride(N, S, true)
  constexpr int REC = (3 * N + 1 + S) * NP;         // G | W | m | b | K
  constexpr int TILE = lds_bwd_packed_tile_doubles(N, S);
  constexpr int SA = S > 0 ? S : 1;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x;
  const int T = p.T;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  double* tile0 = smem + 2;
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (lane >= ACT) return;
  const int sub = lane / N, c = lane - sub * N;
  const long long seq = (long long)blockIdx.x * SPW + sub;
  const bool valid = seq < p.B;
  const long long b = valid ? seq : (long long)p.B - 1;

  double* tile = tile0 + (size_t)sub * TILE;
  double* ring = tile;                       // 2 x record prefix
  double* SbS = ring + 2 * REC;              // SbS[k*NP+i] = Sb[i][k]
  double* CbS = SbS + N * NP;                // CbS[k*NP+i] = Cb[i][k]; then Gb, then (sweep 2) Pb, same indexing
  double* vbS = CbS + N * NP;                // vbS[v*NP+i]: vector cotangent v (0 = mu, 1+s = x^s), element i
  double* epS = vbS + (1 + S) * NP;          // eps_t[s][i]
  double* wm = epS + S * NP;                 // W mb      (sweep 2: hb_{t+1})
  double* us = wm + NP;                      // u = K xb_s

  double* stash = p.stash + (size_t)b * T * RS;
  const size_t bt0 = (size_t)b * T;
  const double* gEx = p.g_E_node_x ? p.g_E_node_x + bt0 * N + c : nullptr;
  const double* gExx = p.g_E_node_xx ? p.g_E_node_xx + bt0 * N + c : nullptr;
  const double* gS = (S > 0 && p.g_samples) ? p.g_samples + bt0 * S * N + c : nullptr;
  const double* xS = S > 0 ? p.samples + bt0 * S * N + c : nullptr;
  const double* eS = S > 0 ? p.eps + bt0 * S * N + c : nullptr;
  const double* Ex = p.E_node_x + bt0 * N + c;
  const double g_ln = p.g_lognorm ? p.g_lognorm[b] : 0.0;

  // one lane issues the record copies of all the warp's sequences (uniform datapath), as in the packed forward kernel
  auto fetch = [&](int t, int slot) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[slot], (uint32_t)(SPW * REC * sizeof(double)));
#pragma unroll
      for (int s2 = 0; s2 < SPW; ++s2) {
        long long bs = (long long)blockIdx.x * SPW + s2;
        if (bs >= p.B) bs = (long long)p.B - 1;
        tma_load_1d(tile0 + s2 * TILE + slot * REC, p.stash + ((size_t)bs * T + t) * RS, (uint32_t)(REC * sizeof(double)),
                    &bars[slot]);
      }
    }
  };
  fetch(0, 0);
  if (T > 1) fetch(1, 1);

  // per-lane operands of a step, straight from global memory
  double sig[N];                              // Sigma_{t+1}[:, c]
  auto load_sig = [&](int t) {                // for step t (t + 1 < T)
    const double* sp = stash + (size_t)(t + 1) * RS + (3 * N + 1 + S) * NP + c;
#pragma unroll
    for (int i = 0; i < N; ++i) sig[i] = sp[i * NP];
  };
  double gx = 0.0, gxx = 0.0, mut = 0.0, mun = 0.0, gs[SA], xn[SA], ep[SA];
#pragma unroll
  for (int s = 0; s < SA; ++s) gs[s] = xn[s] = ep[s] = 0.0;
  auto load_in = [&](int t) {
    gx = gEx ? gEx[(size_t)t * N] : 0.0;
    gxx = gExx ? gExx[(size_t)t * N] : 0.0;
    mut = Ex[(size_t)t * N];
    if (t + 1 < T) mun = Ex[(size_t)(t + 1) * N];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      gs[s] = gS ? gS[((size_t)t * S + s) * N] : 0.0;
      ep[s] = eS[((size_t)t * S + s) * N];
      if (t + 1 < T) xn[s] = xS[((size_t)(t + 1) * S + s) * N];
    }
  };
  if (T > 1) load_sig(0);
  load_in(0);

  double Sb[N];                               // column c of the cotangent of Sigma_t
  double vb[1 + SA];                          // element c of the cotangents of mu_t, x^s_t
#pragma unroll
  for (int i = 0; i < N; ++i) Sb[i] = 0.0;
#pragma unroll
  for (int v = 0; v < 1 + SA; ++v) vb[v] = 0.0;

  // ================================ sweep 1: forward in time ==================================
  for (int t = 0; t < T; ++t) {
    const bool last = (t == T - 1);
    const double* rec = ring + (t & 1) * REC;
    const double* G = rec;                    // G[k*NP+i] = G_t[i][k]
    const double* W = rec + N * NP;
    const double* mrow = rec + 2 * N * NP;
    const double* K = rec + (2 * N + 1 + S) * NP;

    // ---- seeds ---------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i == c) Sb[i] += gxx;
    vb[0] += fma(2.0 * gxx, mut, gx);          // E[x x']_cc = Sigma_cc + mu_c^2: its cotangent reaches mu_c too
#pragma unroll
    for (int s = 0; s < S; ++s) vb[1 + s] += gs[s];
#pragma unroll
    for (int q2 = 0; q2 < NP / 2; ++q2) st2(&SbS[c * NP + 2 * q2], Sb[2 * q2], (2 * q2 + 1 < N) ? Sb[2 * q2 + 1] : 0.0);
#pragma unroll
    for (int v = 0; v < 1 + S; ++v) vbS[v * NP + c] = vb[v];
#pragma unroll
    for (int s = 0; s < S; ++s) epS[s * NP + c] = ep[s];
    const double mun_c = mun;
    double xn_c[SA];
#pragma unroll
    for (int s = 0; s < SA; ++s) xn_c[s] = xn[s];
    mbar_wait(&bars[t & 1], (uint32_t)((t >> 1) & 1));
    __syncwarp(MASK);

    // Register discipline (168 registers at 9..12 resident warps per SM): once published to shared memory, Sb and vb are
    // dead, so the carries G' Cb and G' vb are written straight into them; Cb lives only from its product to the carry;
    // (Sb W) joins `inner` late, from the same broadcast as (Gb G').
    if (!last) {
      double grow[N];                         // G_t[k][c] over k: row c of the transposed storage
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        const double2 w = ld2(&G[c * NP + 2 * q2]);
        grow[2 * q2] = w.x;
        if (2 * q2 + 1 < N) grow[2 * q2 + 1] = w.y;
      }
      // vb_{t+1}[c] = (G' vb)[c]
#pragma unroll
      for (int v = 0; v < 1 + S; ++v) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 w = ld2(&vbS[v * NP + 2 * q2]);
          s0 = fma(grow[2 * q2], w.x, s0);
          if (2 * q2 + 1 < N) s1 = fma(grow[2 * q2 + 1], w.y, s1);
        }
        vb[v] = s0 + s1;
      }
      {
        // Cb[:,c] = Sb G[:,c]
        double Cb[N];
#pragma unroll
        for (int i = 0; i < N; ++i) Cb[i] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 v = ld2(&SbS[k * NP + 2 * q2]);
            Cb[2 * q2] = fma(v.x, grow[k], Cb[2 * q2]);
            if (2 * q2 + 1 < N) Cb[2 * q2 + 1] = fma(v.y, grow[k], Cb[2 * q2 + 1]);
          }
        }
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) st2(&CbS[c * NP + 2 * q2], Cb[2 * q2], (2 * q2 + 1 < N) ? Cb[2 * q2 + 1] : 0.0);
        // Sb_{t+1}[:,c] = (G' Cb)[:,c]
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 g = ld2(&G[i * NP + 2 * q2]);          // G_t[k][i], k = 2q2, 2q2+1
            s0 = fma(g.x, Cb[2 * q2], s0);
            if (2 * q2 + 1 < N) s1 = fma(g.y, Cb[2 * q2 + 1], s1);
          }
          Sb[i] = s0 + s1;
        }
      }
      double Gb[N];
      {
        // C[:,c] = G Sigma_{t+1}[:,c]
        double C[N];
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
        // Gb[:,c] = Sb' C[:,c] + sum_v vb_v v_{t+1}[c]
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 v = ld2(&SbS[i * NP + 2 * q2]);     // Sb[k][i], k = 2q2, 2q2+1
            a0 = fma(v.x, C[2 * q2], a0);
            if (2 * q2 + 1 < N) a1 = fma(v.y, C[2 * q2 + 1], a1);
          }
          Gb[i] = a0 + a1;
        }
      }
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        const double2 v = ld2(&vbS[2 * q2]);
        Gb[2 * q2] = fma(v.x, mun_c, Gb[2 * q2]);
        if (2 * q2 + 1 < N) Gb[2 * q2 + 1] = fma(v.y, mun_c, Gb[2 * q2 + 1]);
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 v = ld2(&vbS[(1 + s) * NP + 2 * q2]);
          Gb[2 * q2] = fma(v.x, xn_c[s], Gb[2 * q2]);
          if (2 * q2 + 1 < N) Gb[2 * q2 + 1] = fma(v.y, xn_c[s], Gb[2 * q2 + 1]);
        }
      }
      __syncwarp(MASK);                       // CbS visible
#pragma unroll
      for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 v = ld2(&CbS[k * NP + 2 * q2]);
          Gb[2 * q2] = fma(v.x, sig[k], Gb[2 * q2]);           // Cb[:,k] Sigma_{t+1}[c][k]
          if (2 * q2 + 1 < N) Gb[2 * q2 + 1] = fma(v.y, sig[k], Gb[2 * q2 + 1]);
        }
      }
      if (t + 2 < T) load_sig(t + 1);         // sig is dead: next step's column lands during the rest of this step
      __syncwarp(MASK);                       // every lane is done with Cb: the tile now holds Gb
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) st2(&CbS[c * NP + 2 * q2], Gb[2 * q2], (2 * q2 + 1 < N) ? Gb[2 * q2 + 1] : 0.0);
    }
    load_in(t + 1 < T ? t + 1 : t);           // next step's per-lane cotangent elements (issued before the barrier below)
    __syncwarp(MASK);                         // Gb visible
    double wrow[N];                           // W[c][:] (= W[:, c])
#pragma unroll
    for (int q2 = 0; q2 < NP / 2; ++q2) {
      const double2 w = ld2(&W[c * NP + 2 * q2]);
      wrow[2 * q2] = w.x;
      if (2 * q2 + 1 < N) wrow[2 * q2 + 1] = w.y;
    }
    // inner[:,c] = (Sb W)[:,c] + (Gb G')[:,c] = sum_k Sb[:,k] W[k][c] + Gb[:,k] G_t[c][k]
    double inner[N];
#pragma unroll
    for (int i = 0; i < N; ++i) inner[i] = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        const double2 v = ld2(&SbS[k * NP + 2 * q2]);
        inner[2 * q2] = fma(v.x, wrow[k], inner[2 * q2]);
        if (2 * q2 + 1 < N) inner[2 * q2 + 1] = fma(v.y, wrow[k], inner[2 * q2 + 1]);
      }
    }
    if (!last) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double gck = G[k * NP + c];
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 v = ld2(&CbS[k * NP + 2 * q2]);
          inner[2 * q2] = fma(v.x, gck, inner[2 * q2]);
          if (2 * q2 + 1 < N) inner[2 * q2 + 1] = fma(v.y, gck, inner[2 * q2 + 1]);
        }
      }
    }

    // (W mb)_c,  mb = mub + sum_s xb_s
    {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        double2 m2 = ld2(&vbS[2 * q2]);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const double2 a = ld2(&vbS[(1 + s) * NP + 2 * q2]);
          m2.x += a.x;
          m2.y += a.y;
        }
        s0 = fma(wrow[2 * q2], m2.x, s0);
        if (2 * q2 + 1 < N) s1 = fma(wrow[2 * q2 + 1], m2.y, s1);
      }
      wm[c] = s0 + s1;
    }
    // Q[:,c] = -(W inner)[:,c]
    double Q[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        const double2 w = ld2(&W[i * NP + 2 * q2]);
        a0 = fma(w.x, inner[2 * q2], a0);
        if (2 * q2 + 1 < N) a1 = fma(w.y, inner[2 * q2 + 1], a1);
      }
      Q[i] = -(a0 + a1);
    }
    __syncwarp(MASK);                         // wm visible
    const double m_c = mrow[c];
#pragma unroll
    for (int q2 = 0; q2 < NP / 2; ++q2) {
      const double2 w = ld2(&wm[2 * q2]);
      const double2 mm = ld2(&mrow[2 * q2]);
      Q[2 * q2] = fma(-w.x, m_c, Q[2 * q2]);                                       // -(W mb) m'
      Q[2 * q2] = fma(-0.5 * g_ln, fma(mm.x, m_c, wrow[2 * q2]), Q[2 * q2]);       // -g/2 (W + m m')
      if (2 * q2 + 1 < N) {
        Q[2 * q2 + 1] = fma(-w.y, m_c, Q[2 * q2 + 1]);
        Q[2 * q2 + 1] = fma(-0.5 * g_ln, fma(mm.y, m_c, wrow[2 * q2 + 1]), Q[2 * q2 + 1]);
      }
    }
    const double f_c = fma(g_ln, m_c, wm[c]);
    // noise adjoint: sum_s K' Phi(-eps_s u_s') K,  u_s = K xb_s   (K = L^-1 lower triangular)
#pragma unroll
    for (int s = 0; s < S; ++s) {
      {
        double u0 = 0.0, u1 = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 kk = ld2(&K[c * NP + 2 * q2]);         // K[c][k]: zeros above the diagonal
          const double2 xb = ld2(&vbS[(1 + s) * NP + 2 * q2]);
          u0 = fma(kk.x, xb.x, u0);
          if (2 * q2 + 1 < N) u1 = fma(kk.y, xb.y, u1);
        }
        us[c] = u0 + u1;
      }
      __syncwarp(MASK);
      // y_i = -eps_i ( sum_{j<i} u_j K[j][c] + u_i K[i][c] / 2 ),  then  Q[:,c] += K' y
      double run = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double kic = K[i * NP + c];
        const double uk = us[i] * kic;
        const double y = -epS[s * NP + i] * fma(0.5, uk, run);
        run += uk;
#pragma unroll
        for (int q2 = 0; q2 < (i + 2) / 2; ++q2) {
          const double2 kk = ld2(&K[i * NP + 2 * q2]);         // K[i][j], j <= i non-zero
          Q[2 * q2] = fma(kk.x, y, Q[2 * q2]);
          if (2 * q2 + 1 <= i) Q[2 * q2 + 1] = fma(kk.y, y, Q[2 * q2 + 1]);
        }
      }
      __syncwarp(MASK);
    }
    // ---- write Q_t (over the K block) and f_t (over the first row of W) ---------------------------
    if (valid) {
      double* kt = stash + (size_t)t * RS + (2 * N + 1 + S) * NP;
#pragma unroll
      for (int i = 0; i < N; ++i) kt[i * NP + c] = Q[i];
      stash[(size_t)t * RS + N * NP + c] = f_c;
    }
    __syncwarp(MASK);                         // every lane is done with the ring slot, SbS, vbS, epS, wm, us
    if (t + 2 < T) fetch(t + 2, t & 1);
  }

  // ================================ sweep 2: backward in time =================================
  __threadfence_block();                      // Q_t / f_t went through the generic proxy
  fence_proxy_async();
  __syncwarp(MASK);
  // barrier phases continue from sweep 1: slot 0 completed ceil(T/2) times, slot 1 floor(T/2) times
  fetch(T - 1, 0);
  if (T > 1) fetch(T - 2, 1);
  double* PbS = CbS;                          // PbS[k*NP+i] = Pb_{t+1}[i][k]
  double* hbS = wm;                           // hb_{t+1}
  double Pb[N];
#pragma unroll
  for (int i = 0; i < N; ++i) Pb[i] = 0.0;
  double* oJ = p.g_node_J + bt0 * N + c;
  double* oh = p.g_node_h + bt0 * N + c;
  for (int t = T - 1; t >= 0; --t) {
    const int j = T - 1 - t;
    const double* rec = ring + (j & 1) * REC;
    const double* G = rec;
    const double* mrow = rec + 2 * N * NP;
    const double* frow = rec + N * NP;
    const double* Qm = rec + (2 * N + 1 + S) * NP;
    const int used0 = (j & 1) == 0 ? (T + 1) / 2 : T / 2;
    mbar_wait(&bars[j & 1], (uint32_t)((used0 + (j >> 1)) & 1));
    double Pn[N];
#pragma unroll
    for (int i = 0; i < N; ++i) Pn[i] = Qm[i * NP + c];
    double hn_c = frow[c];
    if (t < T - 1) {
      // T1[:,c] = (Pb_{t+1} G')[:,c] = sum_k Pb[:,k] G_t[c][k]
      double T1[N];
#pragma unroll
      for (int i = 0; i < N; ++i) T1[i] = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double gck = G[k * NP + c];
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 v = ld2(&PbS[k * NP + 2 * q2]);
          T1[2 * q2] = fma(v.x, gck, T1[2 * q2]);
          if (2 * q2 + 1 < N) T1[2 * q2 + 1] = fma(v.y, gck, T1[2 * q2 + 1]);
        }
      }
      const double m_c = mrow[c];
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
        Pn[i] += gt1[i] - ghb[i] * m_c;       // Q + G Pb G' - (G hb) m'
        if (i == c) hn_c += ghb[i];           // f + G hb
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) Pb[i] = Pn[i];
    double diag = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i == c) diag = Pb[i];
    if (valid) {
      oJ[(size_t)t * N] = -2.0 * diag;
      oh[(size_t)t * N] = hn_c;
    }
    __syncwarp(MASK);                         // everyone is done with PbS / hbS of t+1 and with the ring slot
#pragma unroll
    for (int q2 = 0; q2 < NP / 2; ++q2) st2(&PbS[c * NP + 2 * q2], Pb[2 * q2], (2 * q2 + 1 < N) ? Pb[2 * q2 + 1] : 0.0);
    hbS[c] = hn_c;
    if (t >= 2) fetch(t - 2, j & 1);
    __syncwarp(MASK);
  }
}

}  // namespace svae
