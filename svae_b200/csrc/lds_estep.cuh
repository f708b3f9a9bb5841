// lds_estep.cuh -- fused LDS local E-step for sm_100a: information-form Kalman filter (forward in
// time) + covariance-form RTS smoother + reparameterised backward sampler (backward in time), for a
// batch of independent sequences, fp64.
//
// What it replaces (reference, one sequence per call, LAPACK on n x n tiles):
//   svae/lds/cython_lds_inference.pyx:28-90    natural_filter_forward_general
//   svae/lds/cython_lds_inference.pyx:149-210  natural_smoother_general + _compute_stats
//   svae/lds/cython_lds_inference.pyx:310-355  natural_sample_backward
//   per-step kernels svae/lds/cython_gaussian_grads.pxd:38-82,125-144,167-188,208-296,431-508
//
// B200 design (not a translation):
//   * one sub-warp tile of LPS lanes per sequence, each lane owning CPL "column slots" (CPL = 2, LPS = 8:
//     FOUR sequences per warp at n <= 14).  Slot c < N is COLUMN c of every n x n matrix, held in
//     registers; slot N is the mean/potential vector, slots N+1..N+S one sample path each.  All register
//     arrays are fully unrolled over the compile-time N; cross-lane traffic is 16-byte broadcast reads from
//     shared memory, and with CPL = 2 every broadcast element feeds two multiply-adds per lane.
//   * filter step t (same arithmetic result as _natural_condition_diag + _natural_predict):
//       M_t = J_pred + diag(J_node) + J11        (information form)
//       M_t = L~ D L~'                           square-root-free LDL' (N reciprocals, no sqrt chain)
//       [Y | m] = M_t^-1 [J12 | h_f]             augmented two-sided solve, lane = rhs column,
//       M_t^-1                                   second rhs set (identity) fused in the same sweep
//       J_pred' = J22 - J12' Y, h_pred' = -J12' m
//       lognorm += 1/2 h_f' m - 1/2 sum log d_k
//     and, off the recursion's critical path, everything the backward sweep needs:
//       G_t = -Y (smoother gain), M_t^-1, m_t, b_t = m_t + L^-T eps_t   -> "stash" record in HBM.
//     Node potentials and noise for step t+1 are staged by cp.async (LDGSTS) while step t computes.
//   * backward step t is pure multiply-add on the stash record, which a 1-D TMA bulk copy
//     (cp.async.bulk + mbarrier, double-buffered) lands in shared memory two steps ahead:
//       [C | mu_t | x_t] = G_t [Sigma_{t+1} | mu_{t+1} | x_{t+1}] + [0 | m_t | b_t]
//       Sigma_t = M_t^-1 + C G_t'
//       E[x_t x_{t+1}'] = C + mu_t mu_{t+1}',  E[x_t x_t'] = Sigma_t + mu_t mu_t'
//     which equals the reference's information-form _rts_backward_step (pxd:240-296) in exact
//     arithmetic and is a sum of PSD terms numerically (no J22 - temp'temp cancellation).
//   * expected sufficient statistics are reduced over time in registers and written once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "async_copy.cuh"

#ifndef SVAE_LDS_TMA
#define SVAE_LDS_TMA 1
#endif
#define SVAE_LDS_CONST_NN 256

namespace svae {

struct LdsFwdParams {
  int B, T, S, flags, const_slot;
  const double *init_J, *init_h, *init_logZ;
  const double *J11, *J12, *J22, *pair_logZ;
  long long pair_stride_t, pair_stride_b;
  const double *node_J, *node_h, *node_logZ;
  const double* eps;
  double *E_init_xx, *E_init_x, *E_pair_xx, *E_pair_xxn, *E_pair_xnxn;
  double *E_node_xx, *E_node_x, *samples, *lognorm, *Exx_t, *Exxn_t;
  double *msg_J_pred, *msg_h_pred;          // optional outputs: prediction messages, natural form (B,T,n,n), (B,T,n)
  const double *in_J_filt, *in_h_filt;      // optional inputs: run the backward sweeps from GIVEN filtered messages
  const double *node_J_dense;               // optional input: dense node precisions (B,T,n,n), natural form (added to node_J)
  const double *term_J;                     // optional input: dense precision (B,n,n) added at the LAST step only
  const int* active;                        // optional (B,) mask: sequences with active[b] == 0 are left untouched
  int* status;
  double* stash;
};

__host__ __device__ constexpr int lds_np(int N) { return (N + 1) & ~1; }   // even row stride

// doubles per (sequence, time) stash record: G (N x NP) | Minv (N x NP) | m (NP) | b (S x NP): this
// prefix is what the smoother sweep fetches.  With SVAE_LDS_SAVE_FOR_BWD the record grows by
// K = L^-1 (N x NP, lower triangular) and the smoothed covariance Sigma_t (N x NP) for the VJP kernel.
__host__ __device__ inline int lds_stash_stride(int N, int S) { return (2 * N + 1 + S) * lds_np(N); }
__host__ __device__ inline int lds_record_stride(int N, int S, bool save) {
  return (save ? 4 * N + 1 + S : 2 * N + 1 + S) * lds_np(N);
}
#define SVAE_LDS_FLAG_SAVE 2
#define SVAE_LDS_FLAG_INIT_PER_SEQ 4

// shared-memory doubles per sequence tile.  The filter sweep (LDL' factors, node ring, per-tile pair
// parameters) and the backward sweep (stash ring, C) never overlap in time, so they share storage.
__host__ __device__ inline int lds_tile_filter_doubles(int N, int S, bool inhomog) {
  const int NP = lds_np(N);
  return 2 * N * NP /*Lr, Lt*/ + 2 * (2 + S) * NP /*node ring*/ + (inhomog ? 5 * N * NP /*tJ12, tJ22, pair stage*/ : 0);
}
__host__ __device__ inline int lds_tile_backward_doubles(int N, int S) {
  return 2 * lds_stash_stride(N, S) /*stash ring*/ + N * lds_np(N) /*Cs*/;
}
__host__ __device__ inline int lds_tile_doubles(int N, int S, bool inhomog) {
  const int NP = lds_np(N);
  const int fa = lds_tile_filter_doubles(N, S, inhomog), fb = lds_tile_backward_doubles(N, S);
  return 2 /*mbarriers*/ + 3 * NP /*sq, mbuf, mu_s*/ + S * NP /*xs*/ + (fa > fb ? fa : fb);
}

// homogeneous J12 (natural form, row-major n x n) lives in the constant bank so the prediction's
// multiply-adds take it as an immediate-offset constant operand: no shared-memory traffic at all.
// (a small pool of slots so launches on different streams do not serialise on one bank)
#define SVAE_LDS_CONST_SLOTS 8
static __constant__ double c_J12[SVAE_LDS_CONST_SLOTS][SVAE_LDS_CONST_NN];

// register budget (one warp per CTA).  CPL = 1, n <= 10: 128 registers so that 16 CTAs fit per SM (the whole
// metric batch in one wave); CPL = 2 keeps two columns per lane in registers and needs the full file, with
// four sequences per warp only 7 CTAs per SM are needed.
#ifdef SVAE_LDS_MINB_OVERRIDE
#define SVAE_LDS_MINB(N, CPL) SVAE_LDS_MINB_OVERRIDE
#endif
#ifndef SVAE_LDS_MINB
#define SVAE_LDS_MINB(N, CPL) (((CPL) == 1 && (N) <= 10) ? 16 : 7)
#endif
#define SVAE_LDS_BOUNDS(N, CPL) __launch_bounds__(32, SVAE_LDS_MINB(N, CPL))

template <int N, int CPL, int LPS, bool INHOMOG>
__global__ void SVAE_LDS_BOUNDS(N, CPL) lds_estep_fwd_kernel(const LdsFwdParams p) {
  constexpr int SPW = 32 / LPS;
  constexpr int NP = lds_np(N);
  constexpr int NN = N * N;
  static_assert(CPL == 1 || CPL == 2, "one or two column slots per lane");
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPS, l = lane % LPS;
  const int S = p.S, T = p.T;
  const long long seq = ((long long)blockIdx.x * nwarps + warp) * SPW + sub;
  const long long b = seq < p.B ? seq : (long long)p.B - 1;
  const bool valid = seq < p.B && (!p.active || p.active[b] != 0);
  // the column slots of this lane
  int c[CPL], sx[CPL], cm[CPL];
  bool mat[CPL], vec[CPL], smp[CPL];
#pragma unroll
  for (int cl = 0; cl < CPL; ++cl) {
    c[cl] = l * CPL + cl;
    mat[cl] = c[cl] < N;
    vec[cl] = c[cl] == N;
    smp[cl] = (c[cl] > N) && (c[cl] <= N + S);
    sx[cl] = smp[cl] ? c[cl] - N - 1 : 0;
    cm[cl] = mat[cl] ? c[cl] : 0;   // safe column index for slots that do not own a matrix column
  }
  const bool anymat = mat[0];                     // slots are ordered: slot 0 is a matrix column iff any is
  const bool pairmat = mat[CPL - 1];              // every slot of this lane is a matrix column
  const int SS = lds_stash_stride(N, S);               // fetched prefix of a record
  const bool save = (p.flags & SVAE_LDS_FLAG_SAVE) != 0;
  const int RS = lds_record_stride(N, S, save);         // record stride in HBM
  const int NODE = (2 + S) * NP;

  // ---- shared memory carve-up (all offsets even => 16-byte aligned) --------------------------
  double* shJ11 = smem;           // CTA-shared homogeneous pair parameters, information form
  double* shJ12 = shJ11 + N * NP;
  double* shJ22 = shJ12 + N * NP;
  double* tile = shJ22 + N * NP + (size_t)(warp * SPW + sub) * lds_tile_doubles(N, S, INHOMOG);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tile);   // 2 mbarriers
  double* sq = tile + 2;          // 1/sqrt(d_k)
  double* mbuf = sq + NP;         // m_t
  double* mu_s = mbuf + NP;       // mu_t
  double* xs = mu_s + NP;         // x_t[s][k]
  double* uni = xs + S * NP;      // ---- phase-shared storage ----
  // filter sweep
  double* Lr = uni;               // Lr[k*NP+i]: raw LDL' column k (i >= k), Lr[k*NP+k] = d_k - node
  double* Lt = Lr + N * NP;       // Lt[i*NP+k]: unit-lower L~[i][k] (k < i), row-major
  double* nring = Lt + N * NP;    // 2 x { J_node row | h_node row | eps rows }
  double* tJ12 = nring + 2 * NODE;  // per-tile pair parameters (INHOMOG only)
  double* tJ22 = tJ12 + N * NP;
  double* pstage = tJ22 + N * NP;   // raw J11 | J12 | J22 of the NEXT step, prefetched with cp.async (INHOMOG only)
  // backward sweep
  double* ring = uni;             // 2 stash records
  double* Cs = ring + 2 * SS;     // Cs[k*NP+i] = C[i][k]
  const double* J12s = INHOMOG ? tJ12 : shJ12;
  const double* J22s = INHOMOG ? tJ22 : shJ22;

  if (!INHOMOG) {
    // shJ11 holds J22 + J11 (the information matrix every non-terminal prediction lands in),
    // shJ12 = J12, shJ22 = J22, all in information form
    for (int e = threadIdx.x; e < NN; e += blockDim.x) {
      const int i = e / N, j = e % N;
      shJ11[i * NP + j] = -2.0 * p.J22[e] - 2.0 * p.J11[e];
      shJ12[i * NP + j] = -p.J12[e];
      shJ22[i * NP + j] = -2.0 * p.J22[e];
    }
  }
  if (l == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  double* stash = p.stash + (size_t)b * T * RS;
  const double* nJ = p.node_J + (size_t)b * T * N;
  const double* nh = p.node_h + (size_t)b * T * N;
  const double* eps = p.eps ? p.eps + (size_t)b * T * S * N : nullptr;

  // stage the node potentials (+ noise) of time t into node-ring slot t&1 with cp.async
  auto stage_node = [&](int t) {
#pragma unroll
    for (int cl = 0; cl < CPL; ++cl) {
      if (mat[cl]) {
        double* dst = nring + (t & 1) * NODE;
        cp_async8(dst + c[cl], nJ + (size_t)t * N + c[cl]);
        cp_async8(dst + NP + c[cl], nh + (size_t)t * N + c[cl]);
        for (int ss = 0; ss < S; ++ss)
          cp_async8(dst + (2 + ss) * NP + c[cl], eps + ((size_t)t * S + ss) * N + c[cl]);
      }
    }
    cp_async_commit();
  };

  // INHOMOG: the pair parameters of step t stream into `pstage` one step ahead; every lane copies and later reads
  // only its own column, so the staging needs no barrier of its own (the group is retired by the step's
  // cp_async_wait<0>).  Without it a short batch pays a global-memory round trip at the top of every step.
  auto stage_pair = [&](int t) {
    if (INHOMOG) {
      const size_t off = (size_t)t * p.pair_stride_t + (size_t)b * p.pair_stride_b;
#pragma unroll
      for (int cl = 0; cl < CPL; ++cl) {
        if (mat[cl]) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            cp_async8(pstage + i * NP + c[cl], p.J11 + off + i * N + c[cl]);
            cp_async8(pstage + (N + i) * NP + c[cl], p.J12 + off + i * N + c[cl]);
            cp_async8(pstage + (2 * N + i) * NP + c[cl], p.J22 + off + i * N + c[cl]);
          }
        }
      }
      cp_async_commit();
    }
  };

  double lacc = 0.0;   // per-lane partial of lognorm (reduced over the tile at the end)
  double lq[CPL];      // sum_t sum_k z_k^2 / d_k accumulated by every slot, only the vec slot's is used
#pragma unroll
  for (int cl = 0; cl < CPL; ++cl) lq[cl] = 0.0;
  int badbits = 0;     // OR of the pivots' high words: the sign bit flags a non-positive-definite M_t
  double dprod = 1.0;  // running product of this lane's pivots (log taken every few steps)

  if (T > 1) stage_pair(0);
  stage_node(0);
  if (T > 1) stage_node(1); else cp_async_commit();
  cp_async_wait<1>();
  __syncwarp();

  // a: mat slots -- column c of M_t = J_pred + J11 (the node precision is added at the pivot);
  //    vec slot  -- h_filtered = h_pred + h_node.   Information form.
  double a[CPL][N];
  const bool init_per_seq = (p.flags & SVAE_LDS_FLAG_INIT_PER_SEQ) != 0;
  const double* iJ = p.init_J + (init_per_seq ? (size_t)b * NN : 0);
  const double* ih = p.init_h + (init_per_seq ? (size_t)b * N : 0);
#pragma unroll
  for (int cl = 0; cl < CPL; ++cl) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (mat[cl]) {
        a[cl][i] = -2.0 * iJ[i * N + c[cl]];
        if (!INHOMOG && T > 1) a[cl][i] -= 2.0 * p.J11[i * N + c[cl]];
      } else {
        a[cl][i] = vec[cl] ? ih[i] + nring[NP + i] : 0.0;
      }
    }
  }

  // =========================== forward in time: information filter ===========================
  for (int t = 0; t < T; ++t) {
    const bool last = (t == T - 1);
    if (INHOMOG && !last) {
#pragma unroll
      for (int cl = 0; cl < CPL; ++cl) {
        if (mat[cl]) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            a[cl][i] += -2.0 * pstage[i * NP + c[cl]];
            tJ12[i * NP + c[cl]] = -pstage[(N + i) * NP + c[cl]];
            tJ22[i * NP + c[cl]] = -2.0 * pstage[(2 * N + i) * NP + c[cl]];
          }
        }
      }
      __syncwarp();
    }
    const double* nd = nring + (t & 1) * NODE;

    if (p.msg_J_pred || p.in_J_filt || p.node_J_dense || (p.term_J && last)) {
      // cold path: the three-primitive interface of the reference (messages out / messages in)
      const size_t bt = (size_t)b * T + t;
#pragma unroll
      for (int cl = 0; cl < CPL; ++cl) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          if (mat[cl]) {
            const double j11 = last ? 0.0 : (INHOMOG ? -2.0 * pstage[i * NP + c[cl]]   // still this step's J11
                                                     : shJ11[i * NP + c[cl]] - shJ22[i * NP + c[cl]]);
            if (p.msg_J_pred && valid) p.msg_J_pred[bt * NN + i * N + c[cl]] = -0.5 * (a[cl][i] - j11);
            if (p.in_J_filt) a[cl][i] = -2.0 * p.in_J_filt[bt * NN + i * N + c[cl]] + j11;
            if (p.node_J_dense) a[cl][i] -= 2.0 * p.node_J_dense[bt * NN + i * N + c[cl]];
            if (p.term_J && last) a[cl][i] -= 2.0 * p.term_J[(size_t)b * NN + i * N + c[cl]];
          } else if (vec[cl]) {
            if (p.msg_h_pred && valid) p.msg_h_pred[bt * N + i] = a[cl][i] - nd[NP + i];
            if (p.in_h_filt) a[cl][i] = p.in_h_filt[bt * N + i];
          }
        }
      }
    }

    // the staging buffer is consumed: stream the next step's pair parameters into it; they land during this step
    // and are retired by its cp_async_wait<0>
    if (INHOMOG && t + 1 < T - 1) stage_pair(t + 1);

    // right-hand sides: r = J12 column (mat), q = identity column (mat -> M^-1 column)
    double r[CPL][N], q[CPL][N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (CPL == 2) {
        const double2 v = last ? make_double2(0.0, 0.0) : ld2(&J12s[i * NP + cm[0]]);   // cm[0] is even
        r[0][i] = v.x;
        r[CPL - 1][i] = v.y;
      } else {
        r[0][i] = last ? 0.0 : J12s[i * NP + cm[0]];
      }
#pragma unroll
      for (int cl = 0; cl < CPL; ++cl) q[cl][i] = (i == c[cl]) ? 1.0 : 0.0;
    }

    // ---- square-root-free LDL' of M_t fused with the forward substitution of all right-hand
    //      sides: slots = rows; the pivot column is broadcast through Lr; the vec slot's `a`
    //      rides along as a bordered row, which IS its forward substitution ----------------------
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int cl = 0; cl < CPL; ++cl)
        if (mat[cl]) Lr[k * NP + c[cl]] = a[cl][k];
      __syncwarp();
      double row[NP];
#pragma unroll
      for (int q2 = k / 2; q2 < NP / 2; ++q2) {
        const double2 v = ld2(&Lr[k * NP + 2 * q2]);
        row[2 * q2] = v.x;
        row[2 * q2 + 1] = v.y;
      }
      const double piv = fma(-2.0, nd[k], row[k]);     // + node precision on the diagonal
      badbits |= __double2hiint(piv);
      const double ik = fast_rcp(piv);
#pragma unroll
      for (int cl = 0; cl < CPL; ++cl) {
        const double ta = a[cl][k] * ik, tr = r[cl][k] * ik, tq = q[cl][k] * ik;
        if (mat[cl]) Lt[c[cl] * NP + k] = ta;                   // unit-lower factor L~[c][k], row-major
        lq[cl] = fma(a[cl][k], ta, lq[cl]);                     // z_k^2 / d_k (meaningful in the vec slot)
#pragma unroll
        for (int j = k + 1; j < N; ++j) {
          a[cl][j] = fma(-ta, row[j], a[cl][j]);
          r[cl][j] = fma(-tr, row[j], r[cl][j]);
          q[cl][j] = fma(-tq, row[j], q[cl][j]);
        }
        a[cl][k] = ta;
        r[cl][k] = tr;
        q[cl][k] = tq;
      }
    }
#pragma unroll
    for (int cl = 0; cl < CPL; ++cl) {
      if (mat[cl]) {
        const double dc = fma(-2.0, nd[c[cl]], Lr[c[cl] * NP + c[cl]]);
        dprod *= dc;
        const double rs = rsqrt(dc);
        sq[c[cl]] = rs;
        if (save) mbuf[c[cl]] = dc * rs;   // sqrt(d_c), consumed below (mbuf is free until m_t is published)
      }
      if (vec[cl]) {
#pragma unroll
        for (int i = 0; i < N; ++i) r[cl][i] = a[cl][i];
      }
    }
    if ((t & 3) == 3 || last) {
      lacc -= 0.5 * log(dprod);
      dprod = 1.0;
    }
    __syncwarp();
#pragma unroll
    for (int cl = 0; cl < CPL; ++cl) {
      if (smp[cl]) {
        const double* ep = nd + (2 + sx[cl]) * NP;
#pragma unroll
        for (int k = 0; k < N; ++k) r[cl][k] = ep[k] * sq[k];   // D^-1/2 eps
      }
    }
    if (save) {
      // K = L^-1 column c: (L~^-1 e_c)_i / sqrt(d_i) = q_i sqrt(d_i)   (q is D^-1 L~^-1 e_c here)
      double* kt = stash + (size_t)t * RS + (2 * N + 1 + S) * NP;
#pragma unroll
      for (int cl = 0; cl < CPL; ++cl) {
        if (mat[cl] && valid) {
#pragma unroll
          for (int i = 0; i < N; ++i) kt[i * NP + c[cl]] = q[cl][i] * mbuf[i];
        }
      }
      __syncwarp();   // mbuf is rewritten with m_t below
    }
    // ---- backward substitution with the unit-upper factor L~' (in place), column-oriented: once y_i is
    //      final, r[k] -= L~[i][k] y_i for all k < i -- independent accumulators, chain of one FMA per i ----
#pragma unroll
    for (int i = N - 1; i >= 1; --i) {
#pragma unroll
      for (int q2 = 0; q2 < (i + 1) / 2; ++q2) {
        const double2 v = ld2(&Lt[i * NP + 2 * q2]);          // L~[i][2q2], L~[i][2q2+1]
#pragma unroll
        for (int cl = 0; cl < CPL; ++cl) {
          r[cl][2 * q2] = fma(-v.x, r[cl][i], r[cl][2 * q2]);
          q[cl][2 * q2] = fma(-v.x, q[cl][i], q[cl][2 * q2]);
          if (2 * q2 + 1 < i) {
            r[cl][2 * q2 + 1] = fma(-v.y, r[cl][i], r[cl][2 * q2 + 1]);
            q[cl][2 * q2 + 1] = fma(-v.y, q[cl][i], q[cl][2 * q2 + 1]);
          }
        }
      }
    }
    // mat: r = M^-1 J12[:,c] = -G[:,c], q = M^-1[:,c];  vec: r = m;  smp: r = L^-T eps
#pragma unroll
    for (int cl = 0; cl < CPL; ++cl) {
      if (vec[cl]) {
#pragma unroll
        for (int i = 0; i < N; ++i) mbuf[i] = r[cl][i];
      }
    }
    __syncwarp();
#pragma unroll
    for (int cl = 0; cl < CPL; ++cl) {
      if (smp[cl]) {
#pragma unroll
        for (int i = 0; i < N; ++i) r[cl][i] += mbuf[i];   // b_t = m_t + L^-T eps_t
      }
    }

    if (!last || save) {
      // ---- stash record for the backward sweep (and, when saving, for the VJP kernel) ---------
      double* st = stash + (size_t)t * RS;
      if (valid) {
        // G is stored TRANSPOSED (Gt[k*NP+i] = G[i][k]): slot c owns the contiguous row c, and the smoother's
        // C = G z becomes an axpy over k with independent accumulators.  M^-1 is symmetric: same trick.
#pragma unroll
        for (int cl = 0; cl < CPL; ++cl) {
          if (mat[cl]) {
#pragma unroll
            for (int q2 = 0; q2 < NP / 2; ++q2) {
              st2(&st[c[cl] * NP + 2 * q2], -r[cl][2 * q2], (2 * q2 + 1 < N) ? -r[cl][2 * q2 + 1] : 0.0);
              st2(&st[N * NP + c[cl] * NP + 2 * q2], q[cl][2 * q2], (2 * q2 + 1 < N) ? q[cl][2 * q2 + 1] : 0.0);
            }
            st[2 * N * NP + c[cl]] = mbuf[c[cl]];
          } else if (smp[cl]) {
#pragma unroll
            for (int i = 0; i < N; ++i) st[(2 * N + 1 + sx[cl]) * NP + i] = r[cl][i];
          }
        }
      }
    }
    if (!last) {
      // next step's node slot has landed; its ring slot from two steps ago is free again
      cp_async_wait<0>();
      __syncwarp();
      if (t + 2 < T) stage_node(t + 2);
      // ---- predict straight into next step's M / h_f:
      //      [J_pred' + J11 | h_pred' + h_node'] = [J22 + J11 | h_node'] - J12' [Y | m] -----------
      {
        const double* ndn = nring + ((t + 1) & 1) * NODE;
        const double* pm = INHOMOG ? J22s : ((t + 1 == T - 1) ? shJ22 : shJ11);
#pragma unroll
        for (int cl = 0; cl < CPL; ++cl) {
          const double* pinit = mat[cl] ? pm + c[cl] : ndn + NP;
          const int pstr = mat[cl] ? NP : 1;
#pragma unroll
          for (int i = 0; i < N; ++i) a[cl][i] = pinit[i * pstr];
        }
        if (INHOMOG) {
#pragma unroll
          for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int q2 = 0; q2 < NP / 2; ++q2) {
              const double2 v = ld2(&J12s[k * NP + 2 * q2]);
#pragma unroll
              for (int cl = 0; cl < CPL; ++cl) {
                a[cl][2 * q2] = fma(-v.x, r[cl][k], a[cl][2 * q2]);
                if (2 * q2 + 1 < N) a[cl][2 * q2 + 1] = fma(-v.y, r[cl][k], a[cl][2 * q2 + 1]);
              }
            }
          }
        } else {
          // information-form J12 = -(natural J12):  a -= J12_info' r  ==  a += J12_nat' r
          const double* cj12 = c_J12[p.const_slot];
#pragma unroll
          for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
              const double j = cj12[k * N + i];
#pragma unroll
              for (int cl = 0; cl < CPL; ++cl) a[cl][i] = fma(j, r[cl][k], a[cl][i]);
            }
          }
        }
      }
    } else {
      // terminal step: the smoothed marginal at T-1 is the filtered one
#pragma unroll
      for (int cl = 0; cl < CPL; ++cl) {
#pragma unroll
        for (int i = 0; i < N; ++i) a[cl][i] = mat[cl] ? q[cl][i] : r[cl][i];
      }
    }
    __syncwarp();
  }
  // backward-sweep state: columns of [Sigma | mu | x]  (reuses the registers of a)
  double (&z)[CPL][N] = a;

  // lognorm: reduce the per-lane partials over the tile, add the constant terms
  {
    double tot = lacc;
#pragma unroll
    for (int cl = 0; cl < CPL; ++cl)
      if (vec[cl]) tot = fma(0.5, lq[cl], tot);
#pragma unroll
    for (int o = LPS / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o, LPS);
    if (l == 0 && valid && p.lognorm) {
      double konst = p.init_logZ ? p.init_logZ[init_per_seq ? b : 0] : 0.0;
      if (p.pair_logZ) {
        if (!INHOMOG) {
          konst += (double)(T - 1) * p.pair_logZ[0];
        } else {
          const size_t lz0 = (size_t)b * p.pair_stride_b / NN, lzs = (size_t)p.pair_stride_t / NN;   // strides are multiples of NN
          for (int t = 0; t < T - 1; ++t) konst += p.pair_logZ[lz0 + (size_t)t * lzs];
        }
      }
      if (p.node_logZ)
        for (int t = 0; t < T; ++t) konst += p.node_logZ[(size_t)b * T + t];
      p.lognorm[b] = tot + konst;
    }
  }
  // negative pivot (sign bit) or a NaN that reached the normaliser
  if ((badbits < 0 || !(lacc == lacc)) && p.status) atomicMax(p.status, (int)b + 1);

  // =========================== backward in time: smoother + sampler ==========================
  // stash records were written through the generic proxy; order them before the async-proxy reads
  __threadfence_block();
  fence_proxy_async();
  __syncwarp();
  auto fetch = [&](int t, int j) {   // j-th backward step (t = T-2-j) -> ring slot j&1
#if SVAE_LDS_TMA
    if (l == 0) {
      mbar_arrive_expect_tx(&bars[j & 1], (uint32_t)(SS * sizeof(double)));
      tma_load_1d(ring + (j & 1) * SS, stash + (size_t)t * RS, (uint32_t)(SS * sizeof(double)), &bars[j & 1]);
    }
#else
    double* dst = ring + (j & 1) * SS;
    const double* src = stash + (size_t)t * RS;
    for (int e = l; e < SS; e += LPS) dst[e] = src[e];
#endif
  };
  if (T >= 2) fetch(T - 2, 0);
  if (T >= 3) fetch(T - 3, 1);

  double acc_mid[CPL][N], acc_xxn[CPL][N];
  double mun_c[CPL];
#pragma unroll
  for (int cl = 0; cl < CPL; ++cl) {
    mun_c[cl] = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) acc_mid[cl][i] = acc_xxn[cl][i] = 0.0;
  }

  for (int t = T - 1; t >= 0; --t) {
    if (t < T - 1) {
      const int j = T - 2 - t;
      const double* rec = ring + (j & 1) * SS;
#if SVAE_LDS_TMA
      mbar_wait(&bars[j & 1], (uint32_t)((j >> 1) & 1));
#else
      __syncwarp();
#endif
      const double* G = rec;
      const double* Minv = rec + N * NP;
      // C = G z  for every column slot of [Sigma_{t+1} | mu_{t+1} | x_{t+1}]:  cc[:] += Gt[k][:] z[k]
      {
        double cc[CPL][N];
#pragma unroll
        for (int cl = 0; cl < CPL; ++cl)
#pragma unroll
          for (int i = 0; i < N; ++i) cc[cl][i] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 g = ld2(&G[k * NP + 2 * q2]);
#pragma unroll
            for (int cl = 0; cl < CPL; ++cl) {
              cc[cl][2 * q2] = fma(g.x, z[cl][k], cc[cl][2 * q2]);
              if (2 * q2 + 1 < N) cc[cl][2 * q2 + 1] = fma(g.y, z[cl][k], cc[cl][2 * q2 + 1]);
            }
          }
        }
#pragma unroll
        for (int cl = 0; cl < CPL; ++cl) {
          if (mat[cl]) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
              Cs[c[cl] * NP + i] = cc[cl][i];
              acc_xxn[cl][i] += cc[cl][i];   // Cov(x_t, x_{t+1})[i][c]; the mean outer product is added below
            }
          } else if (vec[cl] || smp[cl]) {
            const double* base = rec + (2 * N + (vec[cl] ? 0 : 1 + sx[cl])) * NP;
#pragma unroll
            for (int i = 0; i < N; ++i) z[cl][i] = base[i] + cc[cl][i];
          }
        }
      }
      __syncwarp();
      if (anymat) {
        if (CPL == 2) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const double2 v = ld2(&Minv[i * NP + cm[0]]);
            if (mat[0]) z[0][i] = v.x;
            if (mat[CPL - 1]) z[CPL - 1][i] = v.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) z[0][i] = Minv[i * NP + cm[0]];
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
          double gk[CPL];
#pragma unroll
          for (int cl = 0; cl < CPL; ++cl) gk[cl] = G[k * NP + cm[cl]];   // G[c][k], transposed storage
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 v = ld2(&Cs[k * NP + 2 * q2]);
#pragma unroll
            for (int cl = 0; cl < CPL; ++cl) {
              if (mat[cl]) {
                z[cl][2 * q2] = fma(v.x, gk[cl], z[cl][2 * q2]);
                if (2 * q2 + 1 < N) z[cl][2 * q2 + 1] = fma(v.y, gk[cl], z[cl][2 * q2 + 1]);   // Sigma_t[:, c]
              }
            }
          }
        }
      }
    }
    // ---- emit statistics for time t -----------------------------------------------------------
#pragma unroll
    for (int cl = 0; cl < CPL; ++cl) {
      if (vec[cl]) {
#pragma unroll
        for (int i = 0; i < N; ++i) mu_s[i] = z[cl][i];
      } else if (smp[cl]) {
#pragma unroll
        for (int i = 0; i < N; ++i) xs[sx[cl] * NP + i] = z[cl][i];
      }
    }
    __syncwarp();
    const size_t bt = (size_t)b * T + t;
    const bool first = (t == 0), lastt = (t == T - 1);
#pragma unroll
    for (int cl = 0; cl < CPL; ++cl) {
      if (mat[cl] && save && valid) {
        double* sg = stash + (size_t)t * RS + (3 * N + 1 + S) * NP;
#pragma unroll
        for (int i = 0; i < N; ++i) sg[i * NP + c[cl]] = z[cl][i];   // Sigma_t[:, c]
      }
      if (mat[cl]) {
        const int cc_ = c[cl];
        const double mu_c = mu_s[cc_];
        // hot path: running sums only
        double diag = 0.0;
        if (!first && !lastt) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const double mi = mu_s[i];
            const double exx = fma(mi, mu_c, z[cl][i]);           // E[x_t x_t'][i][c]
            if (i == cc_) diag = exx;
            acc_xxn[cl][i] = fma(mi, mun_c[cl], acc_xxn[cl][i]);
            acc_mid[cl][i] += exx;
          }
        } else {
          // first / last time step: E_init, the end corrections of the pair sums
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const double mi = mu_s[i];
            const double exx = fma(mi, mu_c, z[cl][i]);
            if (i == cc_) diag = exx;
            if (!lastt) acc_xxn[cl][i] = fma(mi, mun_c[cl], acc_xxn[cl][i]);
            if (valid) {
              const size_t o = (size_t)b * NN + i * N + cc_;
              if (lastt && p.E_pair_xnxn) p.E_pair_xnxn[o] = (T > 1) ? exx : 0.0;
              if (first) {
                if (p.E_init_xx) p.E_init_xx[o] = exx;
                if (p.E_pair_xx) p.E_pair_xx[o] = (T > 1) ? acc_mid[cl][i] + exx : 0.0;
                if (p.E_pair_xxn) p.E_pair_xxn[o] = acc_xxn[cl][i];
                if (p.E_pair_xnxn && T > 1) p.E_pair_xnxn[o] = (lastt ? 0.0 : p.E_pair_xnxn[o]) + acc_mid[cl][i];
              }
            }
          }
          if (first && valid && p.E_init_x) p.E_init_x[(size_t)b * N + cc_] = mu_c;
        }
        if (valid) {
          if (p.E_node_x) p.E_node_x[bt * N + cc_] = mu_c;
          if (p.E_node_xx) p.E_node_xx[bt * N + cc_] = diag;
          if (p.samples)
            for (int ss = 0; ss < S; ++ss) p.samples[(bt * S + ss) * N + cc_] = xs[ss * NP + cc_];
          if (p.Exx_t) {   // per-time second moments (inhomogeneous E_pair): cold path
#pragma unroll
            for (int i = 0; i < N; ++i) {
              p.Exx_t[bt * NN + i * N + cc_] = fma(mu_s[i], mu_c, z[cl][i]);
              if (!lastt && p.Exxn_t)
                p.Exxn_t[((size_t)b * (T - 1) + t) * NN + i * N + cc_] = fma(mu_s[i], mun_c[cl], Cs[cc_ * NP + i]);
            }
          }
        }
        mun_c[cl] = mu_c;
      }
    }
    __syncwarp();
    if (t >= 2 && t < T - 1) fetch(t - 2, T - t);   // the slot just consumed is free: prefetch
  }
}

}  // namespace svae
