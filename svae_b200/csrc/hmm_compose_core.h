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
#define SVAE_UNROLL4 _Pragma("unroll 4")
#else
#define SVAE_UNROLL
#define SVAE_UNROLL4
#endif

namespace svae {

// Tile layout (compile-time strides so every shared-memory address in the hot loop is an immediate):
//   A   [k*32 + j]            exp(log_A - amax)
//   PDt [k*kHmmRow + i]       (P D)^T, double-buffered: row k = column k of P times d_k, contiguous in i so the product
//                             loop reads it with 128-bit broadcasts; row stride 34 keeps rows 16-byte aligned and the
//                             per-lane row stores at a 4-way bank conflict instead of 32-way.
// One CTA of kHmmWarps warps per chunk: thread (w, j) owns rows [8w, 8w+8) of column j of the running product, so a
// step's 32 x 32 x 32 product is spread over four warp schedulers (a single warp is bounded by the FP64 issue rate of
// its scheduler: 1024 DFMA at one per two cycles).
constexpr int kHmmRow = 34;
constexpr int kHmmWarps = 4;
constexpr int kHmmRows = 32 / kHmmWarps;   // rows per thread
SVAE_HD int hmm_compose_tile_doubles(int /*K*/) { return 32 * 32 + 2 * 32 * kHmmRow + 32 * kHmmWarps + 3 * 32; }

// log_A (K,K) row = from-state; node (T,K) of this sequence.  Writes P (K,K) row-major with max entry 1 and
// *logscale such that Psi = P * exp(*logscale).
SVAE_HD void hmm_compose_chunk(double* mem, int K, int s, int L, const double* log_A, const double* node, double* out_P,
                               double* out_logscale) {
  constexpr int TH = 32 * kHmmWarps;
  double* A = mem;
  double* PDt0 = mem + 32 * 32;
  double* PDt1 = PDt0 + 32 * kHmmRow;
  double* tmax = PDt1 + 32 * kHmmRow;   // per-thread maxima (TH)
  double* colmax = tmax + TH;
  double* d = colmax + 32;
  double* nd = d + 32;
  SVAE_BPHASE(j, K) {
    double m = log_A[j];
    for (int k = 1; k < K; ++k) m = fmax(m, log_A[k * K + j]);
    colmax[j] = m;
    nd[j] = node[(size_t)s * K + j];
    if (L > 1) d[j] = node[(size_t)(s + 1) * K + j];
  } SVAE_BPHASE_END
  double amax = colmax[0], nmax = nd[0];
  for (int j = 1; j < K; ++j) {
    amax = fmax(amax, colmax[j]);
    nmax = fmax(nmax, nd[j]);
  }
  // P = I, so (P D_s)^T row k = d_k e_k; columns and rows beyond K stay exactly zero
  SVAE_BPHASE(t, TH) {
    const int w = t >> 5, j = t & 31;
    if (w == 0 && j < K) {
      for (int k = 0; k < K; ++k) A[k * 32 + j] = exp(log_A[k * K + j] - amax);
    }
    const double dj = (j < K) ? exp(nd[j] - nmax) : 0.0;
    SVAE_UNROLL
    for (int ii = 0; ii < kHmmRows; ++ii) {
      const int i = w * kHmmRows + ii;
      PDt0[j * kHmmRow + i] = (i == j) ? dj : 0.0;
      PDt1[j * kHmmRow + i] = 0.0;
    }
  } SVAE_BPHASE_END
  // node potentials ride a two-slot ring: slot (q+1)&1 holds step s+q+1 at iteration q (slot 0 = nd, slot 1 = d)
  double logscale = (double)L * amax + nmax;
  double inv = 1.0;                      // 1 / max of the previous (unscaled) product, folded into the A operand
  for (int q = 0; q < L; ++q) {
    const double* cur = (q & 1) ? PDt1 : PDt0;
    double* nxt = (q & 1) ? PDt0 : PDt1;
    const bool more = q + 1 < L;
    const double* ndc = ((q + 1) & 1) ? d : nd;     // node potentials of step s+q+1
    double* ndn = ((q + 1) & 1) ? nd : d;           // free slot (held step s+q): receives step s+q+2
    double nmax2 = 0.0;
    if (more) {
      nmax2 = ndc[0];
      for (int j = 1; j < K; ++j) nmax2 = fmax(nmax2, ndc[j]);
    }
    SVAE_BPHASE(t, TH) {
      const int w = t >> 5, j = t & 31;
      if (w == 0 && j < K && q + 2 < L) {  // streams in under the product; retired in the reduction phase below
        SVAE_COPY_ASYNC8(&ndn[j], &node[(size_t)(s + q + 2) * K + j]);
        SVAE_COPY_COMMIT();
      }
      double acc[kHmmRows];                // fixed trip counts keep it in registers on the device
      SVAE_UNROLL
      for (int ii = 0; ii < kHmmRows; ++ii) acc[ii] = 0.0;
      if (j < K) {
        SVAE_UNROLL4                       // overlap the next rows' loads with this row's multiply-adds
        for (int k = 0; k < K; ++k) {
          const double a = A[k * 32 + j] * inv;
          const double* row = cur + k * kHmmRow + w * kHmmRows;
          SVAE_UNROLL
          for (int ii = 0; ii < kHmmRows; ++ii) acc[ii] = fma(row[ii], a, acc[ii]);   // padding rows are zeros
        }
      }
      double m = 0.0;
      const double dj = (j < K) ? (more ? exp(ndc[j] - nmax2) : 1.0) : 0.0;
      SVAE_UNROLL
      for (int ii = 0; ii < kHmmRows; ++ii) {
        m = fmax(m, acc[ii]);
        nxt[j * kHmmRow + w * kHmmRows + ii] = acc[ii] * dj;   // row j of the next (P D)^T (or of P^T after the last step)
      }
      tmax[t] = m;
    } SVAE_BPHASE_END
    SVAE_BPHASE(j, 32) {
      double m = tmax[j];
      SVAE_UNROLL
      for (int w = 1; w < kHmmWarps; ++w) m = fmax(m, tmax[w * 32 + j]);
      colmax[j] = m;
      SVAE_COPY_WAIT_ALL();
    } SVAE_BPHASE_END
    double pmax = colmax[0];
    for (int j = 1; j < 32; ++j) pmax = fmax(pmax, colmax[j]);
    inv = 1.0 / pmax;
    logscale += log(pmax) + (more ? nmax2 : 0.0);
  }
  const double* fin = (L & 1) ? PDt1 : PDt0;
  SVAE_BPHASE(j, K) {
    for (int i = 0; i < K; ++i) out_P[i * K + j] = fin[j * kHmmRow + i] * inv;
    if (j == 0) *out_logscale = logscale;
  } SVAE_BPHASE_END
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
      double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;   // four chains: the step is latency-, not throughput-bound
      int i = 0;
      for (; i + 3 < K; i += 4) {
        p0 = fma(a[i], cur[i * kChainRow + j], p0);
        p1 = fma(a[i + 1], cur[(i + 1) * kChainRow + j], p1);
        p2 = fma(a[i + 2], cur[(i + 2) * kChainRow + j], p2);
        p3 = fma(a[i + 3], cur[(i + 3) * kChainRow + j], p3);
      }
      for (; i < K; ++i) p0 = fma(a[i], cur[i * kChainRow + j], p0);
      tmp[j] = (p0 + p1) + (p2 + p3);
    } SVAE_PHASE_END
    {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int j = 0;
      for (; j + 3 < K; j += 4) { s0 += tmp[j]; s1 += tmp[j + 1]; s2 += tmp[j + 2]; s3 += tmp[j + 3]; }
      for (; j < K; ++j) s0 += tmp[j];
      sum = (s0 + s1) + (s2 + s3);
    }
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
      double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
      int j = 0;
      for (; j + 3 < K; j += 4) {
        p0 = fma(cur[i * kChainRow + j], b[j], p0);
        p1 = fma(cur[i * kChainRow + j + 1], b[j + 1], p1);
        p2 = fma(cur[i * kChainRow + j + 2], b[j + 2], p2);
        p3 = fma(cur[i * kChainRow + j + 3], b[j + 3], p3);
      }
      for (; j < K; ++j) p0 = fma(cur[i * kChainRow + j], b[j], p0);
      tmp[i] = (p0 + p1) + (p2 + p3);
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

// ---------------------------------------------------------------------------------------------------------------
// Max-plus twins of the two stages above, for the most likely state path (svae/hmm/hmm_inference.py:54-63):
// the chunk element is V(i, j) = max over interior paths of the summed log-potentials from z_s = i to z_{s+L} = j
// (node potentials s .. s+L-1 included), the boundary chain runs the max-product recursion over the chunk matrices
// and back-tracks the optimal state at every chunk boundary.  Stage 3 is svae_hmm_viterbi on all chunks with both end
// states pinned (svae_b200/hmm/hmm_pit.py).  Validated on CPU through the host build (tests/test_hmm_pit_cpu.py).
// ---------------------------------------------------------------------------------------------------------------
namespace svae {

constexpr double kHmmNever = -1e300;   // "impossible", finite so that sums of two of them stay finite

SVAE_HD int hmm_maxplus_tile_doubles(int /*K*/) { return 32 * 32 + 2 * 32 * kHmmRow + 32; }

SVAE_HD void hmm_compose_chunk_maxplus(double* mem, int K, int s, int L, const double* log_A, const double* node,
                                       double* out_V) {
  constexpr int TH = 32 * kHmmWarps;
  double* A = mem;                       // log A, element (k, j) at [k*32 + j]
  double* Vt0 = mem + 32 * 32;           // V transposed, row k = column k of V, contiguous in i
  double* Vt1 = Vt0 + 32 * kHmmRow;
  double* nd = Vt1 + 32 * kHmmRow;       // node potentials of the current step
  SVAE_BPHASE(t, TH) {
    const int w = t >> 5, j = t & 31;
    if (w == 0 && j < K) {
      for (int k = 0; k < K; ++k) A[k * 32 + j] = log_A[k * K + j];
      nd[j] = node[(size_t)s * K + j];
    }
    SVAE_UNROLL
    for (int ii = 0; ii < kHmmRows; ++ii) {
      const int i = w * kHmmRows + ii;
      Vt0[j * kHmmRow + i] = (i == j && j < K) ? 0.0 : kHmmNever;   // max-plus identity
      Vt1[j * kHmmRow + i] = kHmmNever;
    }
  } SVAE_BPHASE_END
  for (int q = 0; q < L; ++q) {
    const double* cur = (q & 1) ? Vt1 : Vt0;
    double* nxt = (q & 1) ? Vt0 : Vt1;
    SVAE_BPHASE(t, TH) {
      const int w = t >> 5, j = t & 31;
      double acc[kHmmRows];
      SVAE_UNROLL
      for (int ii = 0; ii < kHmmRows; ++ii) acc[ii] = kHmmNever;
      if (j < K) {
        SVAE_UNROLL4
        for (int k = 0; k < K; ++k) {
          const double a = nd[k] + A[k * 32 + j];          // leave state k at this step, enter j at the next
          const double* row = cur + k * kHmmRow + w * kHmmRows;
          SVAE_UNROLL
          for (int ii = 0; ii < kHmmRows; ++ii) acc[ii] = fmax(acc[ii], row[ii] + a);
        }
      }
      SVAE_UNROLL
      for (int ii = 0; ii < kHmmRows; ++ii) nxt[j * kHmmRow + w * kHmmRows + ii] = acc[ii];
    } SVAE_BPHASE_END
    if (q + 1 < L) {
      SVAE_BPHASE(j, K) { nd[j] = node[(size_t)(s + q + 1) * K + j]; } SVAE_BPHASE_END
    }
  }
  const double* fin = (L & 1) ? Vt1 : Vt0;
  SVAE_BPHASE(j, K) {
    for (int i = 0; i < K; ++i) out_V[i * K + j] = fin[j * kHmmRow + i];
  } SVAE_BPHASE_END
}

SVAE_HD int hmm_chain_viterbi_tile_doubles(int /*K*/) { return 3 * 32; }

// Boundary chain, max-product: delta (C+1, K) scratch receives the forward max-messages (delta[c][j] = best score of
// any path ending in z_{s_c} = j, node potential at s_c excluded); zb (C+1) receives the optimal boundary states
// (first maximal index on ties, like numpy.argmax).
SVAE_HD void hmm_chain_viterbi(double* mem, int K, int C, const double* log_pi, const double* V, const double* node_last,
                               double* delta, int* zb) {
  double* a = mem;
  double* cand = mem + 32;
  double* tmp = mem + 64;
  SVAE_PHASE(j, K) {
    a[j] = log_pi[j];
    delta[j] = a[j];
  } SVAE_PHASE_END
  for (int c = 0; c < C; ++c) {
    const double* Vc = V + (size_t)c * K * K;
    SVAE_PHASE(j, K) {
      double m = kHmmNever * 4.0;
      for (int i = 0; i < K; ++i) m = fmax(m, a[i] + Vc[i * K + j]);
      tmp[j] = m;
    } SVAE_PHASE_END
    SVAE_PHASE(j, K) {
      a[j] = tmp[j];
      delta[(size_t)(c + 1) * K + j] = tmp[j];
    } SVAE_PHASE_END
  }
  SVAE_PHASE(j, K) { cand[j] = a[j] + node_last[j]; } SVAE_PHASE_END
  int z = 0;
  for (int j = 1; j < K; ++j)
    if (cand[j] > cand[z]) z = j;
  SVAE_PHASE(j, 1) { zb[C] = z; } SVAE_PHASE_END
  for (int c = C - 1; c >= 0; --c) {
    const double* Vc = V + (size_t)c * K * K;
    SVAE_PHASE(i, K) { cand[i] = delta[(size_t)c * K + i] + Vc[i * K + z]; } SVAE_PHASE_END
    int zi = 0;
    for (int i = 1; i < K; ++i)
      if (cand[i] > cand[zi]) zi = i;
    z = zi;
    SVAE_PHASE(j, 1) { zb[c] = z; } SVAE_PHASE_END
  }
}

}  // namespace svae
