// hmm_wide_core.h -- HMM local E-step and Viterbi for MORE than 32 states (32 < K <= 1024): the capability the narrow
// register-resident kernel of hmm_estep.cu (lane = state, K <= 32) does not have and the reference has for any K
// (svae/hmm/cython_hmm_inference.pyx:93-166 hmm_logZ / hmm_logZ_grad, svae/hmm/hmm_inference.py:21-63).
//
// One CTA of KP = 32*ceil(K/32) threads per sequence, thread j = state j.  Same scaled recursions as the narrow kernel:
//   forward :  e_t[j] = (sum_i ahat_{t-1}[i] P[i][j]) exp(node_t[j] - max_t),   ahat = e / sum(e),   P = exp(log A - gmax)
//              logZ = sum_t (log sum(e_t) + max_t) + (T-1) gmax;   ahat_t is parked in the E[z_t] output buffer
//   backward:  w_{t+1} = exp(node_{t+1} - max) .* bet_{t+1},   bu = P w_{t+1},   s_t = <ahat_t, bu>
//              E[z_t] = ahat_t .* bu / s_t,   bet_t = bu / s_t,   E[#i->j] = P[i][j] sum_t ahat_t[i] w_{t+1}[j] / s_t
// P (row-major: column access of the forward sweep is coalesced over j) and its transpose (row access of the backward
// sweep, coalesced likewise) are built once per call by hmm_wide_prep into a caller-provided workspace that every CTA reads
// through L1/L2.  The K x K transition-count accumulator lives in the sequence's own block of the Etrans OUTPUT buffer and
// is updated every kHmmWideQ steps from a ring of pending (ahat_t, w_{t+1}/s_t) pairs: one read-modify-write of the block
// per Q steps instead of per step.
//
// Phase discipline of lds_compose_core.h / hmm_compose_core.h: inside a phase thread j writes only entry j (or column j) of
// the tile arrays and reads what EARLIER phases wrote; everything a thread carries across phases lives in the tile.  The
// same source therefore runs on the host with the threads looped (tests/emul/hmm_wide_host.cpp, tests/test_hmm_wide_cpu.py).
#pragma once
#include <math.h>
#include <stdint.h>
#include "lds_compose_core.h"   // SVAE_HD, SVAE_BPHASE, SVAE_BPHASE_END

namespace svae {

constexpr int kHmmWideQ = 8;           // pending outer products per flush of the transition counts
constexpr int kHmmWideMaxK = 1024;     // one thread per state

SVAE_HD int hmm_wide_kp(int K) { return (K + 31) & ~31; }
// e (2 KP, double-buffered message) | x, wv, bu, av, bet (KP each) | aq, wq (Q KP each) | red, redm (32 each) | scal (4)
SVAE_HD int hmm_wide_tile_doubles(int K) { return (7 + 2 * kHmmWideQ) * hmm_wide_kp(K) + 64 + 4; }
SVAE_HD int hmm_wide_prep_tile_doubles(int K) { return hmm_wide_kp(K); }
// workspace doubles: P | P' | gmax
SVAE_HD long long hmm_wide_workspace_doubles(int K) { return 2LL * K * K + 2; }

// P = exp(log_A - max log_A), its transpose, and the shift.  One CTA of KP threads.
SVAE_HD void hmm_wide_prep(double* mem, int K, const double* log_A, double* P, double* PT, double* gmax_out) {
  const int KP = hmm_wide_kp(K);
  double* x = mem;
  SVAE_BPHASE(t, KP) {
    double m = -INFINITY;
    for (int e = t; e < K * K; e += KP) m = fmax(m, log_A[e]);
    x[t] = m;
  } SVAE_BPHASE_END
  SVAE_BPHASE(t, KP) {
    double g = x[0];
    for (int i = 1; i < KP; ++i) g = fmax(g, x[i]);
    for (int e = t; e < K * K; e += KP) {
      const int i = e / K, j = e - i * K;
      const double v = exp(log_A[e] - g);
      P[e] = v;
      PT[(size_t)j * K + i] = v;
    }
    if (t == 0) *gmax_out = g;
  } SVAE_BPHASE_END
}

struct HmmWideTile {
  double *e0, *e1, *x, *wv, *bu, *av, *bet, *aq, *wq, *red, *redm, *scal;
};
SVAE_HD HmmWideTile hmm_wide_tile(double* mem, int K) {
  const int KP = hmm_wide_kp(K);
  HmmWideTile t;
  t.e0 = mem; t.e1 = mem + KP; t.x = mem + 2 * KP; t.wv = mem + 3 * KP; t.bu = mem + 4 * KP; t.av = mem + 5 * KP;
  t.bet = mem + 6 * KP; t.aq = mem + 7 * KP; t.wq = t.aq + kHmmWideQ * KP; t.red = t.wq + kHmmWideQ * KP;
  t.redm = t.red + 32; t.scal = t.redm + 32;
  return t;
}

// per-warp partial sum of v[0..KP) (entries >= K are kept at 0) and partial max of u[0..KP) (entries >= K at -inf):
// every lane of warp w reads the warp's 32 entries (broadcast reads, same order in every lane), lane 0 publishes them
#define SVAE_HMM_WIDE_PARTIALS(v, u, KP_, red_, redm_)                                          \
  SVAE_BPHASE(t_, KP_) {                                                                         \
    if ((t_ & 31) == 0) {                                                                         \
      double s_ = 0.0, m_ = -INFINITY;                                                            \
      for (int l_ = 0; l_ < 32; ++l_) { s_ += (v)[t_ + l_]; m_ = fmax(m_, (u)[t_ + l_]); }        \
      (red_)[t_ >> 5] = s_; (redm_)[t_ >> 5] = m_;                                                \
    }                                                                                             \
  } SVAE_BPHASE_END

// One sequence.  node (T,K); P, PT (K,K) from hmm_wide_prep; logZ_out (1) or null; Ez (T,K) (required unless fwd_only);
// Ez0 (K) or null; Etrans (K,K) or null.  The CTA-uniform flag fwd_only stops after the log-normaliser.
SVAE_HD void hmm_wide_sequence(double* mem, int K, int T, const double* log_pi, const double* P, const double* PT, double gmax,
                               const double* node, double* logZ_out, double* Ez0, double* Etrans, double* Ez, bool fwd_only) {
  const int KP = hmm_wide_kp(K), NW = KP / 32;
  const HmmWideTile tl = hmm_wide_tile(mem, K);
  double* red = tl.red; double* redm = tl.redm; double* scal = tl.scal;
  double* x = tl.x;
  const double NEG = -INFINITY;

  // ------------------------------------------------ forward ------------------------------------------------
  SVAE_BPHASE(j, KP) {
    x[j] = (j < K) ? log_pi[j] + node[j] : NEG;
    tl.e0[j] = 0.0;                           // so that the partial sums below see zeros past K
    tl.e1[j] = 0.0;
    if (j == 0) scal[0] = 0.0;
  } SVAE_BPHASE_END
  SVAE_HMM_WIDE_PARTIALS(tl.e0, x, KP, red, redm)
  SVAE_BPHASE(j, KP) {
    double m = redm[0];
    for (int w = 1; w < NW; ++w) m = fmax(m, redm[w]);
    if (j < K) tl.e0[j] = exp(x[j] - m);
    if (j == 0) scal[0] = m;
  } SVAE_BPHASE_END
  for (int t = 1; t < T; ++t) {
    double* ep = (t & 1) ? tl.e0 : tl.e1;     // message t-1 (unnormalised)
    double* en = (t & 1) ? tl.e1 : tl.e0;
    SVAE_BPHASE(j, KP) { x[j] = (j < K) ? node[(size_t)t * K + j] : NEG; } SVAE_BPHASE_END
    SVAE_HMM_WIDE_PARTIALS(ep, x, KP, red, redm)
    SVAE_BPHASE(j, KP) {
      double c = red[0], m = redm[0];
      for (int w = 1; w < NW; ++w) { c += red[w]; m = fmax(m, redm[w]); }
      const double invc = 1.0 / c;
      if (j < K) {
        double p0 = 0.0, p1 = 0.0;
        int i = 0;
        for (; i + 1 < K; i += 2) {
          p0 = fma(ep[i], P[(size_t)i * K + j], p0);
          p1 = fma(ep[i + 1], P[(size_t)(i + 1) * K + j], p1);
        }
        if (i < K) p0 = fma(ep[i], P[(size_t)i * K + j], p0);
        en[j] = (p0 + p1) * invc * exp(x[j] - m);
        if (!fwd_only) Ez[(size_t)(t - 1) * K + j] = ep[j] * invc;      // ahat_{t-1}
      }
      if (j == 0) scal[0] += log(c) + m;
    } SVAE_BPHASE_END
  }
  {
    double* el = ((T - 1) & 1) ? tl.e1 : tl.e0;
    SVAE_BPHASE(j, KP) { x[j] = NEG; } SVAE_BPHASE_END
    SVAE_HMM_WIDE_PARTIALS(el, x, KP, red, redm)
    SVAE_BPHASE(j, KP) {
      double c = red[0];
      for (int w = 1; w < NW; ++w) c += red[w];
      if (j < K && !fwd_only) Ez[(size_t)(T - 1) * K + j] = el[j] / c;     // ahat_{T-1} = E[z_{T-1}]
      if (j == 0) {
        const double lz = scal[0] + log(c) + (double)(T - 1) * gmax;
        if (logZ_out) *logZ_out = lz;
      }
    } SVAE_BPHASE_END
  }
  if (fwd_only) return;

  // ------------------------------------------------ backward -----------------------------------------------
  SVAE_BPHASE(j, KP) {
    tl.bet[j] = (j < K) ? 1.0 : 0.0;
    tl.wv[j] = 0.0;
    tl.bu[j] = 0.0;
    if (Etrans && j < K)
      for (int i = 0; i < K; ++i) Etrans[(size_t)i * K + j] = 0.0;
  } SVAE_BPHASE_END
  int npend = 0;
  for (int t = T - 2; t >= 0; --t) {
    SVAE_BPHASE(j, KP) { x[j] = (j < K) ? node[(size_t)(t + 1) * K + j] : NEG; } SVAE_BPHASE_END
    SVAE_HMM_WIDE_PARTIALS(tl.bu, x, KP, red, redm)          // only the max is used here
    SVAE_BPHASE(j, KP) {
      double m = redm[0];
      for (int w = 1; w < NW; ++w) m = fmax(m, redm[w]);
      tl.wv[j] = (j < K) ? exp(x[j] - m) * tl.bet[j] : 0.0;           // w_{t+1}
    } SVAE_BPHASE_END
    SVAE_BPHASE(j, KP) {
      double prd = 0.0;
      if (j < K) {
        double p0 = 0.0, p1 = 0.0;
        int i = 0;
        for (; i + 1 < K; i += 2) {
          p0 = fma(tl.wv[i], PT[(size_t)i * K + j], p0);
          p1 = fma(tl.wv[i + 1], PT[(size_t)(i + 1) * K + j], p1);
        }
        if (i < K) p0 = fma(tl.wv[i], PT[(size_t)i * K + j], p0);
        const double b = p0 + p1;                                      // (P w_{t+1})[j]
        const double a = Ez[(size_t)t * K + j];                        // ahat_t (this thread stored it)
        tl.bu[j] = b;
        tl.av[j] = a;
        prd = a * b;
      }
      x[j] = prd;      // x is free again: the per-warp sums below run over it
    } SVAE_BPHASE_END
    SVAE_HMM_WIDE_PARTIALS(x, tl.bu, KP, red, redm)          // only the sum is used here
    SVAE_BPHASE(j, KP) {
      double s = red[0];
      for (int w = 1; w < NW; ++w) s += red[w];
      const double invs = 1.0 / s;
      if (j < K) {
        Ez[(size_t)t * K + j] = x[j] * invs;
        tl.bet[j] = tl.bu[j] * invs;
      }
      tl.aq[npend * KP + j] = (j < K) ? tl.av[j] : 0.0;
      tl.wq[npend * KP + j] = tl.wv[j] * invs;
    } SVAE_BPHASE_END
    ++npend;
    if (npend == kHmmWideQ || t == 0) {
      if (Etrans) {
        SVAE_BPHASE(j, KP) {
          if (j < K) {
            for (int i = 0; i < K; ++i) {
              double acc = Etrans[(size_t)i * K + j];
              for (int q = 0; q < npend; ++q) acc = fma(tl.aq[q * KP + i], tl.wq[q * KP + j], acc);
              Etrans[(size_t)i * K + j] = acc;
            }
          }
        } SVAE_BPHASE_END
      }
      npend = 0;
    }
  }
  SVAE_BPHASE(j, KP) {
    if (j < K) {
      if (Ez0) Ez0[j] = Ez[j];
      if (Etrans)
        for (int i = 0; i < K; ++i) Etrans[(size_t)i * K + j] *= P[(size_t)i * K + j];
    }
  } SVAE_BPHASE_END
}

// Viterbi, one sequence: max-product forward with 16-bit back-pointers (T,K) in the workspace, serial back-track by
// thread 0.  First maximal index on ties (NumPy argmax), like the narrow kernel.  Tile: 2 KP doubles.
SVAE_HD int hmm_wide_viterbi_tile_doubles(int K) { return 2 * hmm_wide_kp(K); }
SVAE_HD void hmm_wide_viterbi(double* mem, int K, int T, const double* log_pi, const double* log_A, const double* node,
                              int32_t* path, uint16_t* back) {
  const int KP = hmm_wide_kp(K);
  double* d0 = mem; double* d1 = mem + KP;
  SVAE_BPHASE(j, KP) { d0[j] = (j < K) ? log_pi[j] + node[j] : -INFINITY; d1[j] = -INFINITY; } SVAE_BPHASE_END
  for (int t = 1; t < T; ++t) {
    const double* dp = (t & 1) ? d0 : d1;
    double* dn = (t & 1) ? d1 : d0;
    SVAE_BPHASE(j, KP) {
      if (j < K) {
        double best = -INFINITY;
        int arg = 0;
        for (int i = 0; i < K; ++i) {
          const double v = dp[i] + log_A[(size_t)i * K + j];
          if (v > best) { best = v; arg = i; }      // strict '>' keeps the FIRST maximal index
        }
        dn[j] = best + node[(size_t)t * K + j];
        back[(size_t)t * K + j] = (uint16_t)arg;
      }
    } SVAE_BPHASE_END
  }
  const double* dl = ((T - 1) & 1) ? d1 : d0;
  SVAE_BPHASE(j, KP) {
    if (j == 0) {
      int z = 0;
      double best = dl[0];
      for (int i = 1; i < K; ++i)
        if (dl[i] > best) { best = dl[i]; z = i; }
      path[T - 1] = z;
      for (int t = T - 2; t >= 0; --t) {
        z = back[(size_t)(t + 1) * K + z];
        path[t] = z;
      }
    }
  } SVAE_BPHASE_END
}

}  // namespace svae
