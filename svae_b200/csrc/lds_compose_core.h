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
SVAE_HD int compose_tile_doubles(int n) { const int NP = compose_np(n); return 9 * n * NP + 8 * NP; }
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
  double *e1, *e2, *u, *mult, *dinv, *nJ, *nh, *acc;   // vectors (NP each)
};

SVAE_HD ComposeTile compose_tile(double* mem, int n) {
  const int NP = compose_np(n), MM = n * NP;
  ComposeTile t;
  t.M = mem; t.R = mem + MM; t.W = mem + 2 * MM; t.L11 = mem + 3 * MM; t.L12T = mem + 4 * MM; t.L22 = mem + 5 * MM;
  t.J11 = mem + 6 * MM; t.J12 = mem + 7 * MM; t.J22 = mem + 8 * MM;
  double* v = mem + 9 * MM;
  t.e1 = v; t.e2 = v + NP; t.u = v + 2 * NP; t.mult = v + 3 * NP; t.dinv = v + 4 * NP; t.nJ = v + 5 * NP; t.nh = v + 6 * NP;
  t.acc = v + 7 * NP;
  return t;
}

// Compose steps [s, s+L) of sequence-local arrays node_J/node_h (T,n) natural form (diag -1/2 precision), node_logZ
// (T) or null; pair natural form (J11,J12,J22 (n,n) row-major, logZ).  Writes the chunk element in NATURAL form, the
// layout the fused kernel takes as per-(b,t) pair parameters: out_J11 = -1/2 L11, out_J12 = -L12, out_J22 = -1/2 L22
// (n,n row-major), out_logZ = c, out_e1, out_e2 (n).  Returns (uniformly) 0, or 1 + the step of a non-positive pivot.
// Time-inhomogeneous pair parameters: pair_stride_t doubles between consecutive steps' (n,n) blocks (0 = homogeneous);
// pJ11/pJ12/pJ22 point at this sequence's step 0 and pair_logZ (or null) at its step-0 entry with stride
// pair_stride_t / (n*n), exactly the addressing of the fused kernel (lds_estep.cuh).
SVAE_HD int compose_chunk(double* mem, int n, int s, int L, const double* pJ11, const double* pJ12, const double* pJ22,
                          const double* pair_logZ, long long pair_stride_t, const double* node_J, const double* node_h,
                          const double* node_logZ, double* out_J11, double* out_J12, double* out_J22, double* out_logZ,
                          double* out_e1, double* out_e2) {
  constexpr int NP = 16;
  ComposeTile t = compose_tile(mem, n);
  double cc = 0.0;
  int bad = 0;
  const long long lz_stride = pair_stride_t / ((long long)n * n);
  SVAE_PHASE(c, n) {
    const size_t po = (size_t)s * pair_stride_t;
    for (int i = 0; i < n; ++i) {
      t.J11[i * NP + c] = -2.0 * pJ11[po + i * n + c];
      t.J12[i * NP + c] = -pJ12[po + i * n + c];
      t.J22[i * NP + c] = -2.0 * pJ22[po + i * n + c];
    }
  } SVAE_PHASE_END
  // first step: the element of node_s . pair
  SVAE_PHASE(c, n) {
    for (int i = 0; i < n; ++i) {
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
      SVAE_PHASE(c, n) {
        const size_t po = (size_t)(s + q) * pair_stride_t;
        for (int i = 0; i < n; ++i) {
          t.J11[i * NP + c] = -2.0 * pJ11[po + i * n + c];
          t.J12[i * NP + c] = -pJ12[po + i * n + c];
          t.J22[i * NP + c] = -2.0 * pJ22[po + i * n + c];
        }
      } SVAE_PHASE_END
    }
    // Lane roles.  With 3n+1 <= 32 the three n-column blocks of the augmented matrix [M | J12 | L12' | v] get their
    // own lanes (lane = role*n + column, lane 3n = the vector): a third of the work per lane and phase.  Otherwise
    // lane c < n carries column c of all three blocks and lane n the vector.
    const int roles = (3 * n + 1 <= 32) ? 3 : 1;
    const int vlane = roles * n;
    SVAE_PHASE(l, vlane + 1) {
      const int r = l / n, c = l - r * n;
      if (l == vlane) {
        for (int i = 0; i < n; ++i) t.u[i] = t.e2[i] + node_h[o + i];
      } else {
        for (int i = 0; i < n; ++i) {
          if (roles == 1 || r == 0) t.M[i * NP + c] = t.L22[i * NP + c] + t.J11[i * NP + c] + (i == c ? -2.0 * node_J[o + c] : 0.0);
          if (roles == 1 || r == 1) t.R[i * NP + c] = t.J12[i * NP + c];
          if (roles == 1 || r == 2) t.W[i * NP + c] = t.L12T[i * NP + c];
        }
      }
    } SVAE_PHASE_END
    // elimination without pivoting, ONE phase per pivot: the multipliers of pivot k are row k of the (symmetric)
    // trailing block, left in the tile by the previous phase, times 1/pivot computed by every lane for itself.
    // Afterwards R = L^-1 J12, W = L^-1 L12', u = L^-1 v, pivots D.
    double det = 1.0, logdet = 0.0;      // one log per step, not per pivot (rescued if the product leaves range)
    for (int k = 0; k < n; ++k) {
      const double piv = t.M[k * NP + k];
      if (!(piv > 0.0) && !bad) bad = s + q + 1;
      const double pinv = 1.0 / piv;
      det *= piv;
      if (!(det > 1e-200 && det < 1e200)) { logdet += log(det); det = 1.0; }
      SVAE_PHASE(l, vlane + 1) {
        const int r = l / n, c = l - r * n;
        // loads first, stores last: the tile arrays share one allocation, so interleaved read-modify-writes would
        // serialise on possible aliasing
        double mu[16];
        SVAE_UNROLL
        for (int i = 0; i < 16; ++i) mu[i] = (i > k && i < n) ? t.M[k * NP + i] * pinv : 0.0;
        if (l == vlane) {
          const double uk = t.u[k];
          double vu[16];
          SVAE_UNROLL
          for (int i = 0; i < 16; ++i)
            if (i < n) vu[i] = t.u[i];
          SVAE_UNROLL
          for (int i = 0; i < 16; ++i)
            if (i > k && i < n) t.u[i] = vu[i] - mu[i] * uk;
          t.dinv[k] = pinv;
        } else {
          for (int rr = 0; rr < 3; ++rr) {
            if (roles == 3 && rr != r) continue;
            if (rr == 0 && c <= k) continue;         // the eliminated columns of M are never read again
            double* X = rr == 0 ? t.M : (rr == 1 ? t.R : t.W);
            const double xk = X[k * NP + c];
            double vx[16];
            SVAE_UNROLL
            for (int i = 0; i < 16; ++i)
              if (i < n) vx[i] = X[i * NP + c];
            SVAE_UNROLL
            for (int i = 0; i < 16; ++i)
              if (i > k && i < n) X[i * NP + c] = vx[i] - mu[i] * xk;
          }
        }
      } SVAE_PHASE_END
    }
    double quad = 0.0;
    for (int k = 0; k < n; ++k) quad += t.u[k] * t.u[k] * t.dinv[k];
    logdet += log(det);
    cc += (pair_logZ ? pair_logZ[(size_t)(s + q) * lz_stride] : 0.0) + (node_logZ ? node_logZ[s + q] : 0.0) + 0.5 * quad - 0.5 * logdet;
    // products.  Destinations (L11, L12T, L22, e1, e2) are disjoint from everything read here (W, R, u, dinv, J22).
    SVAE_PHASE(l, vlane + 1) {
      const int r = l / n, c = l - r * n;
      const double* __restrict__ Wt = t.W;
      const double* __restrict__ Rt = t.R;
      double* __restrict__ L11 = t.L11;
      double* __restrict__ L12T = t.L12T;
      double* __restrict__ L22 = t.L22;
      double* __restrict__ e1 = t.e1;
      double* __restrict__ e2 = t.e2;
      const double* __restrict__ J22 = t.J22;
      const bool vec = l == vlane;
      double wd[16], rd[16];              // own column (or u) scaled by D^-1
      SVAE_UNROLL
      for (int k = 0; k < 16; ++k) {
        const double dk = (k < n) ? t.dinv[k] : 0.0;
        wd[k] = (k < n) ? (vec ? t.u[k] : Wt[k * NP + c]) * dk : 0.0;
        rd[k] = (k < n && !vec) ? Rt[k * NP + c] * dk : 0.0;
      }
      const bool do11 = vec || roles == 1 || r == 0;      // W' D^-1 (W col | u)
      const bool do12 = vec || roles == 1 || r == 1;      // R' D^-1 (W col | u)
      const bool do22 = !vec && (roles == 1 || r == 2);   // R' D^-1 (R col)
      for (int i = 0; i < n; ++i) {
        double a11 = 0.0, a12 = 0.0, a22 = 0.0;
        SVAE_UNROLL
        for (int k = 0; k < 16; ++k)
          if (k < n) {
            if (do11) a11 = fma(Wt[k * NP + i], wd[k], a11);
            if (do12) a12 = fma(Rt[k * NP + i], wd[k], a12);
            if (do22) a22 = fma(Rt[k * NP + i], rd[k], a22);
          }
        if (vec) {
          e1[i] -= a11;
          e2[i] = -a12;
        } else {
          if (do11) L11[i * NP + c] -= a11;
          if (do12) L12T[i * NP + c] = -a12;
          if (do22) L22[i * NP + c] = J22[i * NP + c] - a22;
        }
      }
    } SVAE_PHASE_END
  }

  SVAE_PHASE(c, n) {
    for (int i = 0; i < n; ++i) {
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

}  // namespace svae
