// hmm_compose_core.h -- chunk composition and boundary messages for the parallel-in-time HMM E-step
// (the sequential recursion is svae/hmm/hmm_inference.py:44-51 and its reverse sweep
// svae/hmm/cython_hmm_inference.pyx:126-166).
//
// A chunk of L steps [s, s+L) of the chain  exp(node_t(z_t)) A(z_t, z_{t+1})  summed over its interior states is the
// K x K matrix  Psi = prod_t (D_t A),  D_t = diag(exp(node_t)):  Psi(i, j) = weight of all paths from z_s = i to
// z_{s+L} = j (node potentials s .. s+L-1 included, the one at s+L not).  It is kept as P = Psi / scale with
// max(P) = 1 and log(scale) separately.  The boundary chain then needs one vector-matrix product per chunk in each
// direction.
//
// Same phase discipline as lds_compose_core.h: in a phase lane j touches only column j (or entry j) of the tile
// arrays and reads what earlier phases wrote, so a warp runs it with __syncwarp() between phases and a loop over the
// lanes reproduces it on the host (tests/emul).
#pragma once
#include <math.h>
#include "lds_compose_core.h"   // SVAE_HD, SVAE_PHASE, SVAE_PHASE_END

#if defined(__CUDACC__)
#define SVAE_UNROLL _Pragma("unroll")
#else
#define SVAE_UNROLL
#endif

namespace svae {

// Tile layout (compile-time strides so every shared-memory address in the hot loop is an immediate):
//   A   [k*32 + j]            exp(log_A - amax)
//   PDt [k*kHmmRow + i]       (P D)^T, double-buffered: row k = column k of P times d_k, contiguous in i so the product
//                             loop reads it with 128-bit broadcasts; row stride 34 keeps rows 16-byte aligned and the
//                             per-lane row stores at a 4-way bank conflict instead of 32-way.
constexpr int kHmmRow = 34;
SVAE_HD int hmm_compose_tile_doubles(int /*K*/) { return 32 * 32 + 2 * 32 * kHmmRow + 3 * 32; }

// log_A (K,K) row = from-state; node (T,K) of this sequence.  Writes P (K,K) row-major with max entry 1 and
// *logscale such that Psi = P * exp(*logscale).
SVAE_HD void hmm_compose_chunk(double* mem, int K, int s, int L, const double* log_A, const double* node, double* out_P,
                               double* out_logscale) {
  double* A = mem;
  double* PDt0 = mem + 32 * 32;
  double* PDt1 = PDt0 + 32 * kHmmRow;
  double* colmax = PDt1 + 32 * kHmmRow;
  double* d = colmax + 32;
  double* nd = d + 32;
  SVAE_PHASE(j, K) {
    double m = log_A[j];
    for (int k = 1; k < K; ++k) m = fmax(m, log_A[k * K + j]);
    colmax[j] = m;
    nd[j] = node[(size_t)s * K + j];
  } SVAE_PHASE_END
  double amax = colmax[0], nmax = nd[0];
  for (int j = 1; j < K; ++j) {
    amax = fmax(amax, colmax[j]);
    nmax = fmax(nmax, nd[j]);
  }
  // P = I, so (P D_s)^T row k = d_k e_k
  SVAE_PHASE(j, K) {
    for (int k = 0; k < K; ++k) A[k * 32 + j] = exp(log_A[k * K + j] - amax);
    const double dj = exp(nd[j] - nmax);
    for (int i = 0; i < 32; ++i) PDt0[j * kHmmRow + i] = (i == j) ? dj : 0.0;   // entries i >= K stay exactly zero
    if (L > 1) nd[j] = node[(size_t)(s + 1) * K + j];
  } SVAE_PHASE_END
  double logscale = (double)L * amax + nmax;
  double inv = 1.0;                      // 1 / max of the previous (unscaled) product, folded into the A operand
  for (int q = 0; q < L; ++q) {
    const double* cur = (q & 1) ? PDt1 : PDt0;
    double* nxt = (q & 1) ? PDt0 : PDt1;
    const bool more = q + 1 < L;
    double nmax2 = 0.0;
    if (more) {
      nmax2 = nd[0];
      for (int j = 1; j < K; ++j) nmax2 = fmax(nmax2, nd[j]);
    }
    SVAE_PHASE(j, K) {
      double acc[32];                      // fixed trip counts keep it in registers on the device
      SVAE_UNROLL
      for (int i = 0; i < 32; ++i) acc[i] = 0.0;
      for (int k = 0; k < K; ++k) {
        const double a = A[k * 32 + j] * inv;
        const double* row = cur + k * kHmmRow;
        SVAE_UNROLL
        for (int i = 0; i < 32; ++i) acc[i] = fma(row[i], a, acc[i]);   // straight-line: the padding rows are zeros
      }
      double m = 0.0;
      const double dj = more ? exp(nd[j] - nmax2) : 1.0;
      SVAE_UNROLL
      for (int i = 0; i < 32; ++i) {
        m = fmax(m, acc[i]);
        nxt[j * kHmmRow + i] = acc[i] * dj;   // row j of the next (P D)^T (or of P^T after the last step)
      }
      colmax[j] = m;
    } SVAE_PHASE_END
    double pmax = colmax[0];
    for (int j = 1; j < K; ++j) pmax = fmax(pmax, colmax[j]);
    inv = 1.0 / pmax;
    logscale += log(pmax) + (more ? nmax2 : 0.0);
    if (q + 2 < L) {
      SVAE_PHASE(j, K) { nd[j] = node[(size_t)(s + q + 2) * K + j]; } SVAE_PHASE_END
    }
  }
  const double* fin = (L & 1) ? PDt1 : PDt0;
  SVAE_PHASE(j, K) {
    for (int i = 0; i < K; ++i) out_P[i * K + j] = fin[j * kHmmRow + i] * inv;
    if (j == 0) *out_logscale = logscale;
  } SVAE_PHASE_END
}

// chain tile: two K x K matrix buffers (row stride 33: column reads of the forward sweep and row reads of the backward
// sweep are both at most 2-way bank conflicted) + four K-vectors
constexpr int kChainRow = 33;
SVAE_HD int hmm_chain_tile_doubles(int /*K*/) { return 2 * 32 * kChainRow + 4 * 32; }

// Boundary chain of one sequence: C chunk matrices P_c (K,K) with log scales, initial log pi, and the node potential
// of the LAST time step.  Writes, for every boundary c = 0..C:
//   log_alpha[c][j] = log p(z_{s_c} = j, node potentials before s_c) + const  (normalised: logsumexp = 0)
//   log_beta[c][j]  = log p(node potentials from s_c on | z_{s_c} = j) + const (max = 0)
// (both clamped below at kLogFloor so they can be added to node potentials), and the total log-normaliser.
// One warp walks the C matrices twice; the next matrix streams into the other tile buffer (cp.async) while the
// current vector-matrix product runs, so a step costs its ~100 instructions, not a global-memory round trip.
SVAE_HD void hmm_chain_messages(double* mem, int K, int C, const double* log_pi, const double* P, const double* logscale,
                                const double* node_last, double* log_alpha, double* log_beta, double* out_logZ) {
  const double kLogFloor = -700.0;
  double* buf0 = mem;
  double* buf1 = mem + 32 * kChainRow;
  double* a = buf1 + 32 * kChainRow;   // current forward vector (normalised)
  double* b = a + 32;                  // current backward vector
  double* tmp = b + 32;
  double* red = tmp + 32;
  SVAE_PHASE(j, K) {
    tmp[j] = log_pi[j];
    for (int i = 0; i < K; ++i) SVAE_COPY_ASYNC8(&buf0[i * kChainRow + j], &P[(size_t)i * K + j]);
    SVAE_COPY_COMMIT();
  } SVAE_PHASE_END
  double m = tmp[0];
  for (int j = 1; j < K; ++j) m = fmax(m, tmp[j]);
  SVAE_PHASE(j, K) { a[j] = exp(tmp[j] - m); } SVAE_PHASE_END
  double sum = 0.0;
  for (int j = 0; j < K; ++j) sum += a[j];
  double logZ = m + log(sum);
  SVAE_PHASE(j, K) {
    a[j] /= sum;
    log_alpha[j] = fmax(log(a[j]), kLogFloor);
  } SVAE_PHASE_END
  double ls_cur = logscale[0];
  for (int c = 0; c < C; ++c) {
    const double* cur = (c & 1) ? buf1 : buf0;
    double* nxt = (c & 1) ? buf0 : buf1;
    const double ls_next = (c + 1 < C) ? logscale[c + 1] : 0.0;
    SVAE_PHASE(j, K) {
      SVAE_COPY_WAIT_ALL();                       // my share of P_c has landed; the barrier publishes everyone's
    } SVAE_PHASE_END
    SVAE_PHASE(j, K) {
      if (c + 1 < C) {
        const double* Pn = P + (size_t)(c + 1) * K * K;
        for (int i = 0; i < K; ++i) SVAE_COPY_ASYNC8(&nxt[i * kChainRow + j], &Pn[(size_t)i * K + j]);
        SVAE_COPY_COMMIT();
      }
      double acc = 0.0;
      for (int i = 0; i < K; ++i) acc = fma(a[i], cur[i * kChainRow + j], acc);
      tmp[j] = acc;
    } SVAE_PHASE_END
    sum = 0.0;
    for (int j = 0; j < K; ++j) sum += tmp[j];
    logZ += ls_cur + log(sum);
    ls_cur = ls_next;
    SVAE_PHASE(j, K) {
      a[j] = tmp[j] / sum;
      log_alpha[(size_t)(c + 1) * K + j] = fmax(log(a[j]), kLogFloor);
    } SVAE_PHASE_END
  }
  // last node potential closes the normaliser and starts the backward vector
  SVAE_PHASE(j, K) { tmp[j] = node_last[j]; } SVAE_PHASE_END
  m = tmp[0];
  for (int j = 1; j < K; ++j) m = fmax(m, tmp[j]);
  SVAE_PHASE(j, K) {
    b[j] = exp(tmp[j] - m);
    red[j] = a[j] * b[j];
    log_beta[(size_t)C * K + j] = fmax(tmp[j] - m, kLogFloor);
  } SVAE_PHASE_END
  sum = 0.0;
  for (int j = 0; j < K; ++j) sum += red[j];
  logZ += m + log(sum);
  // backward: P_{C-1} is still in its buffer; stream P_{c-1} while P_c is used
  for (int c = C - 1; c >= 0; --c) {
    const double* cur = (c & 1) ? buf1 : buf0;
    double* nxt = (c & 1) ? buf0 : buf1;
    SVAE_PHASE(j, K) {
      SVAE_COPY_WAIT_ALL();
    } SVAE_PHASE_END
    SVAE_PHASE(i, K) {
      if (c > 0) {
        const double* Pn = P + (size_t)(c - 1) * K * K;
        for (int r = 0; r < K; ++r) SVAE_COPY_ASYNC8(&nxt[r * kChainRow + i], &Pn[(size_t)r * K + i]);
        SVAE_COPY_COMMIT();
      }
      double acc = 0.0;
      for (int j = 0; j < K; ++j) acc = fma(cur[i * kChainRow + j], b[j], acc);
      tmp[i] = acc;
    } SVAE_PHASE_END
    m = tmp[0];
    for (int j = 1; j < K; ++j) m = fmax(m, tmp[j]);
    SVAE_PHASE(j, K) {
      b[j] = tmp[j] / m;
      log_beta[(size_t)c * K + j] = fmax(log(b[j]), kLogFloor);
    } SVAE_PHASE_END
  }
  SVAE_PHASE(j, 1) { *out_logZ = logZ; } SVAE_PHASE_END
}

}  // namespace svae
