// lds_compose_core.h -- chunk composition for the parallel-in-time LDS E-step.
//
// A chunk of L consecutive steps of the chain   node_t(x_t) . pair(x_t, x_{t+1}),  t = s .. s+L-1
// (the potentials natural_filter_forward_general eliminates one at a time, svae/lds/lds_inference.py:86-106,
// svae/lds/gaussian.py:46-66) is integrated over its interior states into ONE pair potential on (x_s, x_{s+L}):
//
//   psi(x, y) = exp( -1/2 x'L11 x - x'L12 y - 1/2 y'L22 y + e1'x + e2'y + c ).
//
// Appending step t (M = L22 + diag(Jnode_t) + J11,  v = e2 + hnode_t,  M = L D L'):
//   R = L^-1 J12,  W = L^-1 L12',  u = L^-1 v
//   L11 -= W' D^-1 W      e1 -= W' D^-1 u      L12' <- -R' D^-1 W
//   L22 <- J22 - R' D^-1 R        e2 <- -R' D^-1 u
//   c += logZnode_t + logZpair + 1/2 u' D^-1 u - 1/2 sum log D      (same dropped constants as gaussian.py:52-66)
// i.e. natural_predict on (L22, e2) plus the carry of the dependence on the chunk's first state.
//
// The code is written as PHASES: in each phase lane c touches only column c of the tile matrices and reads what
// earlier phases wrote, so one warp runs it with __syncwarp() around every phase (also BEFORE it: the lane-uniform
// code between phases reads tile memory the next phase may overwrite) and a plain loop over lanes reproduces it on
// the host (tests/emul builds exactly this header with g++ to pin the arithmetic on CPU).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SVAE_HD __host__ __device__ __forceinline__
#else
#define SVAE_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define SVAE_PHASE(c, lanes) { __syncwarp(); const int c = (int)(threadIdx.x & 31); if (c < (lanes))
#define SVAE_PHASE_END __syncwarp(); }
#else
#define SVAE_PHASE(c, lanes) { for (int c = 0; c < (lanes); ++c)
#define SVAE_PHASE_END }
#endif

// Block-wide phases (one CTA of `threads` threads per tile, __syncthreads around each phase): same discipline, for
// tiles whose per-step work is worth more than one warp scheduler.
#if defined(__CUDA_ARCH__)
#define SVAE_BPHASE(t, threads) { __syncthreads(); const int t = (int)threadIdx.x; if (t < (threads))
#define SVAE_BPHASE_END __syncthreads(); }
#else
#define SVAE_BPHASE(t, threads) { for (int t = 0; t < (threads); ++t)
#define SVAE_BPHASE_END }
#endif

// Asynchronous global -> tile copies (prefetch of the next step's operands while this step computes): cp.async on the
// device, a plain copy on the host.  A thread waits for its own copies; the phase barrier publishes them to the warp.
#if defined(__CUDA_ARCH__)
#define SVAE_COPY_ASYNC8(dst, src)                                                                              \
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)),        \
               "l"(__cvta_generic_to_global(src))                                                                \
               : "memory")
#define SVAE_COPY_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define SVAE_COPY_WAIT_ALL() asm volatile("cp.async.wait_group 0;" ::: "memory")
#else
#define SVAE_COPY_ASYNC8(dst, src) (*(dst) = *(src))
#define SVAE_COPY_COMMIT() ((void)0)
#define SVAE_COPY_WAIT_ALL() ((void)0)
#endif

namespace svae {

// tile row stride: a compile-time constant (SVAE_LDS_MAX_N) so that the unrolled loops below address shared memory
// with immediates instead of an integer multiply per access
SVAE_HD int compose_np(int /*n*/) { return 16; }
// doubles of tile memory one chunk needs
SVAE_HD int compose_tile_doubles(int n) { const int NP = compose_np(n); return 9 * n * NP + 4 * NP; }
#if defined(__CUDACC__)
#define SVAE_UNROLL _Pragma("unroll")
#else
#define SVAE_UNROLL
#endif

struct ComposeTile {
  // n x NP matrices, element (i, c) at [i * NP + c]: lane c owns column c
  double *M, *R, *W;          // elimination workspace: M | J12 | L12'
  double *L11, *L12T, *L22;   // running element (L12T = L12 transposed)
  double *J11, *J12, *J22;    // pair potential, information form
  double *e1, *e2, *u, *dinv;   // vectors (NP each)
};

SVAE_HD ComposeTile compose_tile(double* mem, int n) {
  const int NP = compose_np(n), MM = n * NP;
  ComposeTile t;
  t.M = mem; t.R = mem + MM; t.W = mem + 2 * MM; t.L11 = mem + 3 * MM; t.L12T = mem + 4 * MM; t.L22 = mem + 5 * MM;
  t.J11 = mem + 6 * MM; t.J12 = mem + 7 * MM; t.J22 = mem + 8 * MM;
  double* v = mem + 9 * MM;
  t.e1 = v; t.e2 = v + NP; t.u = v + 2 * NP; t.dinv = v + 3 * NP;
  return t;
}

// Compose steps [s, s+L) of sequence-local arrays node_J/node_h (T,n) natural form (diag -1/2 precision), node_logZ
// (T) or null; pair natural form (J11,J12,J22 (n,n) row-major, logZ).  Writes the chunk element in NATURAL form, the
// layout the fused kernel takes as per-(b,t) pair parameters: out_J11 = -1/2 L11, out_J12 = -L12, out_J22 = -1/2 L22
// (n,n row-major), out_logZ = c, out_e1, out_e2 (n).  Returns (uniformly) 0, or 1 + the step of a non-positive pivot.
// Time-inhomogeneous pair parameters: pair_stride_t doubles between consecutive steps' (n,n) blocks (0 = homogeneous);
// pJ11/pJ12/pJ22 point at this sequence's step 0 and pair_logZ (or null) at its step-0 entry with stride
// pair_stride_t / (n*n), exactly the addressing of the fused kernel (lds_estep.cuh).
//
// N is a template parameter like in the fused kernels: every loop below has a static trip count, every tile address
// is an immediate (the first, runtime-n version of this function spent 8 700 warp-instructions per step, most of
// them predicates and address arithmetic; ncu: profiles/r01_lds_compose_summary.txt).
// Lane roles: with 3N+1 <= 32 the three N-column blocks of the augmented matrix [M | J12 | L12' | v] get their own
// lanes (lane = role*N + column, lane 3N = the vector); otherwise lane c < N carries column c of all three blocks
// and lane N the vector.
template <int N>
SVAE_HD int compose_chunk(double* mem, int s, int L, const double* pJ11, const double* pJ12, const double* pJ22,
                          const double* pair_logZ, long long pair_stride_t, const double* node_J, const double* node_h,
                          const double* node_logZ, double* out_J11, double* out_J12, double* out_J22, double* out_logZ,
                          double* out_e1, double* out_e2) {
  constexpr int NP = 16, n = N;
  constexpr int ROLES = (3 * N + 1 <= 32) ? 3 : 1;
  constexpr int VL = ROLES * N;           // the vector lane
  ComposeTile t = compose_tile(mem, n);
  double cc = 0.0;
  int bad = 0;
  const long long lz_stride = pair_stride_t / ((long long)n * n);
  SVAE_PHASE(c, N) {
    const size_t po = (size_t)s * pair_stride_t;
    SVAE_UNROLL
    for (int i = 0; i < N; ++i) {
      t.J11[i * NP + c] = -2.0 * pJ11[po + i * n + c];
      t.J12[i * NP + c] = -pJ12[po + i * n + c];
      t.J22[i * NP + c] = -2.0 * pJ22[po + i * n + c];
    }
  } SVAE_PHASE_END
  // first step: the element of node_s . pair
  SVAE_PHASE(c, N) {
    SVAE_UNROLL
    for (int i = 0; i < N; ++i) {
      t.L11[i * NP + c] = t.J11[i * NP + c] + (i == c ? -2.0 * node_J[(size_t)s * n + c] : 0.0);
      t.L12T[i * NP + c] = t.J12[c * NP + i];
      t.L22[i * NP + c] = t.J22[i * NP + c];
    }
    t.e1[c] = node_h[(size_t)s * n + c];
    t.e2[c] = 0.0;
  } SVAE_PHASE_END
  cc = (pair_logZ ? pair_logZ[(size_t)s * lz_stride] : 0.0) + (node_logZ ? node_logZ[s] : 0.0);

  for (int q = 1; q < L; ++q) {
    const size_t o = (size_t)(s + q) * n;
    if (pair_stride_t != 0) {   // this step's pair potential
      SVAE_PHASE(c, N) {
        const size_t po = (size_t)(s + q) * pair_stride_t;
        SVAE_UNROLL
        for (int i = 0; i < N; ++i) {
          t.J11[i * NP + c] = -2.0 * pJ11[po + i * n + c];
          t.J12[i * NP + c] = -pJ12[po + i * n + c];
          t.J22[i * NP + c] = -2.0 * pJ22[po + i * n + c];
        }
      } SVAE_PHASE_END
    }
    // augmented matrix [M | J12 | L12' | v]
    SVAE_PHASE(l, VL + 1) {
      const int r = l / N, c = l - r * N;
      if (l == VL) {
        SVAE_UNROLL
        for (int i = 0; i < N; ++i) t.u[i] = t.e2[i] + node_h[o + i];
      } else {
        const double nj = (ROLES == 1 || r == 0) ? -2.0 * node_J[o + c] : 0.0;
        SVAE_UNROLL
        for (int i = 0; i < N; ++i) {
          if (ROLES == 1 || r == 0) t.M[i * NP + c] = t.L22[i * NP + c] + t.J11[i * NP + c] + (i == c ? nj : 0.0);
          if (ROLES == 1 || r == 1) t.R[i * NP + c] = t.J12[i * NP + c];
          if (ROLES == 1 || r == 2) t.W[i * NP + c] = t.L12T[i * NP + c];
        }
      }
    } SVAE_PHASE_END
    // elimination without pivoting, ONE phase per pivot: the multipliers of pivot k are row k of the (symmetric)
    // trailing block, left in the tile by the previous phase, times 1/pivot computed by every lane for itself.
    // Afterwards R = L^-1 J12, W = L^-1 L12', u = L^-1 v, pivots D.
    double det = 1.0, logdet = 0.0;      // one log per step, not per pivot (rescued if the product leaves range)
    SVAE_UNROLL
    for (int k = 0; k < N; ++k) {
      const double piv = t.M[k * NP + k];
      if (!(piv > 0.0) && !bad) bad = s + q + 1;
      const double pinv = 1.0 / piv;
      det *= piv;
      if (!(det > 1e-200 && det < 1e200)) { logdet += log(det); det = 1.0; }
      SVAE_PHASE(l, VL + 1) {
        const int r = l / N, c = l - r * N;
        // loads first, stores last: the tile arrays share one allocation, so interleaved read-modify-writes would
        // serialise on possible aliasing
        double mu[N];
        SVAE_UNROLL
        for (int i = k + 1; i < N; ++i) mu[i] = t.M[k * NP + i] * pinv;
        if (l == VL) {
          const double uk = t.u[k];
          double vu[N];
          SVAE_UNROLL
          for (int i = k + 1; i < N; ++i) vu[i] = t.u[i];
          SVAE_UNROLL
          for (int i = k + 1; i < N; ++i) t.u[i] = vu[i] - mu[i] * uk;
          t.dinv[k] = pinv;
        } else {
          SVAE_UNROLL
          for (int rr = 0; rr < (ROLES == 1 ? 3 : 1); ++rr) {
            const int role = (ROLES == 1) ? rr : r;
            if (role == 0 && c <= k) continue;       // the eliminated columns of M are never read again
            double* X = role == 0 ? t.M : (role == 1 ? t.R : t.W);
            const double xk = X[k * NP + c];
            double vx[N];
            SVAE_UNROLL
            for (int i = k + 1; i < N; ++i) vx[i] = X[i * NP + c];
            SVAE_UNROLL
            for (int i = k + 1; i < N; ++i) X[i * NP + c] = vx[i] - mu[i] * xk;
          }
        }
      } SVAE_PHASE_END
    }
    double quad = 0.0;
    SVAE_UNROLL
    for (int k = 0; k < N; ++k) quad += t.u[k] * t.u[k] * t.dinv[k];
    logdet += log(det);
    cc += (pair_logZ ? pair_logZ[(size_t)(s + q) * lz_stride] : 0.0) + (node_logZ ? node_logZ[s + q] : 0.0) + 0.5 * quad - 0.5 * logdet;
    // products  A' D^-1 b  with (A, b, destination) by role:
    //   role 0: W' D^-1 W[:,c]  -> L11[:,c] -=      role 1: R' D^-1 W[:,c] -> L12'[:,c] = -
    //   role 2: R' D^-1 R[:,c]  -> L22[:,c] = J22 - vector lane: W' D^-1 u -> e1 -= ,  R' D^-1 u -> e2 = -
    // Destinations are disjoint from everything read here (W, R, u, dinv, J22): said with __restrict__.
    SVAE_PHASE(l, VL + 1) {
      const int r = l / N, c = l - r * N;
      const double* __restrict__ Wt = t.W;
      const double* __restrict__ Rt = t.R;
      const double* __restrict__ J22 = t.J22;
      double* __restrict__ L11 = t.L11;
      double* __restrict__ L12T = t.L12T;
      double* __restrict__ L22 = t.L22;
      double* __restrict__ e1 = t.e1;
      double* __restrict__ e2 = t.e2;
      const bool vec = l == VL;
      SVAE_UNROLL
      for (int pass = 0; pass < ((ROLES == 1) ? 3 : 2); ++pass) {
        const int role = (ROLES == 1 || vec) ? pass : r;
        if (ROLES == 3 && !vec && pass > 0) continue;
        if (vec && pass > 1) continue;
        // vector lane: pass 0 = (W, u), pass 1 = (R, u)
        const double* A = vec ? (pass == 0 ? Wt : Rt) : (role == 0 ? Wt : Rt);
        const double* bs = vec ? t.u : ((role == 2 ? Rt : Wt) + c);
        const int bstr = vec ? 1 : NP;
        double bd[N];
        SVAE_UNROLL
        for (int k = 0; k < N; ++k) bd[k] = bs[k * bstr] * t.dinv[k];
        double acc[N];
        SVAE_UNROLL
        for (int i = 0; i < N; ++i) acc[i] = 0.0;
        SVAE_UNROLL
        for (int k = 0; k < N; ++k) {
          SVAE_UNROLL
          for (int i = 0; i < N; ++i) acc[i] = fma(A[k * NP + i], bd[k], acc[i]);
        }
        SVAE_UNROLL
        for (int i = 0; i < N; ++i) {
          if (vec) {
            if (pass == 0) e1[i] -= acc[i]; else e2[i] = -acc[i];
          } else if (role == 0) {
            L11[i * NP + c] -= acc[i];
          } else if (role == 1) {
            L12T[i * NP + c] = -acc[i];
          } else {
            L22[i * NP + c] = J22[i * NP + c] - acc[i];
          }
        }
      }
    } SVAE_PHASE_END
  }

  SVAE_PHASE(c, N) {
    SVAE_UNROLL
    for (int i = 0; i < N; ++i) {
      // symmetrise the diagonal blocks against rounding drift; natural-form output
      out_J11[i * n + c] = -0.25 * (t.L11[i * NP + c] + t.L11[c * NP + i]);
      out_J22[i * n + c] = -0.25 * (t.L22[i * NP + c] + t.L22[c * NP + i]);
      out_J12[i * n + c] = -t.L12T[c * NP + i];
    }
    out_e1[c] = t.e1[c];
    out_e2[c] = t.e2[c];
    if (c == 0) *out_logZ = cc;
  } SVAE_PHASE_END
  return bad;
}

// runtime-n dispatch (host harness and kernel)
#define SVAE_COMPOSE_CASE(NN) case NN: return compose_chunk<NN>(mem, s, L, pJ11, pJ12, pJ22, pair_logZ, pair_stride_t, \
                                                                node_J, node_h, node_logZ, out_J11, out_J12, out_J22,   \
                                                                out_logZ, out_e1, out_e2);
SVAE_HD int compose_chunk_n(int n, double* mem, int s, int L, const double* pJ11, const double* pJ12, const double* pJ22,
                            const double* pair_logZ, long long pair_stride_t, const double* node_J, const double* node_h,
                            const double* node_logZ, double* out_J11, double* out_J12, double* out_J22, double* out_logZ,
                            double* out_e1, double* out_e2) {
  switch (n) {
    SVAE_COMPOSE_CASE(1) SVAE_COMPOSE_CASE(2) SVAE_COMPOSE_CASE(3) SVAE_COMPOSE_CASE(4) SVAE_COMPOSE_CASE(5)
    SVAE_COMPOSE_CASE(6) SVAE_COMPOSE_CASE(7) SVAE_COMPOSE_CASE(8) SVAE_COMPOSE_CASE(9) SVAE_COMPOSE_CASE(10)
    SVAE_COMPOSE_CASE(11) SVAE_COMPOSE_CASE(12) SVAE_COMPOSE_CASE(13) SVAE_COMPOSE_CASE(14) SVAE_COMPOSE_CASE(15)
    SVAE_COMPOSE_CASE(16)
    default: return -1;
  }
}
#undef SVAE_COMPOSE_CASE

}  // namespace svae
