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
//   * one sub-warp tile of LPS lanes per sequence (LPS = 16 -> 2 sequences per warp, or 32);
//     lane c < N owns COLUMN c of every n x n matrix in registers, lane N owns the mean/potential
//     vector, lanes N+1..N+S own one sample path each.  All register arrays are fully unrolled over
//     the compile-time N; cross-lane traffic is 16-byte broadcast reads from shared memory.
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
  return 2 * N * NP /*Lr, Lt*/ + 2 * (2 + S) * NP /*node ring*/ + (inhomog ? 2 * N * NP : 0);
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

// register budget: one warp per CTA; SVAE_LDS_MAXREG caps registers/thread (144 -> 14 CTAs per SM)
#if defined(SVAE_LDS_MAXREG)
#define SVAE_LDS_BOUNDS __maxnreg__(SVAE_LDS_MAXREG)
#else
#ifndef SVAE_LDS_MINB
#define SVAE_LDS_MINB 8
#endif
#define SVAE_LDS_BOUNDS __launch_bounds__(32, SVAE_LDS_MINB)
#endif
template <int N, int LPS, bool INHOMOG>
__global__ void SVAE_LDS_BOUNDS lds_estep_fwd_kernel(const LdsFwdParams p) {
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
  const int cm = mat ? c : 0;   // safe column index for lanes that do not own a matrix column
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
  double* Lt = Lr + N * NP;       // Lt[k*NP+i]: unit-lower L~[i][k] (i > k)
  double* nring = Lt + N * NP;    // 2 x { J_node row | h_node row | eps rows }
  double* tJ12 = nring + 2 * NODE;  // per-tile pair parameters (INHOMOG only)
  double* tJ22 = tJ12 + N * NP;
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
  if (c == 0) {
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
    if (mat) {
      double* dst = nring + (t & 1) * NODE;
      cp_async8(dst + c, nJ + (size_t)t * N + c);
      cp_async8(dst + NP + c, nh + (size_t)t * N + c);
      for (int ss = 0; ss < S; ++ss) cp_async8(dst + (2 + ss) * NP + c, eps + ((size_t)t * S + ss) * N + c);
    }
    cp_async_commit();
  };

  double lacc = 0.0;   // per-lane partial of lognorm (reduced over the tile at the end)
  double dprod = 1.0;  // running product of this lane's pivots (log taken every 8 steps)
  bool bad = false;

  stage_node(0);
  if (T > 1) stage_node(1); else cp_async_commit();
  cp_async_wait<1>();
  __syncwarp();

  // a: mat lanes -- column c of M_t = J_pred + J11 (the node precision is added at the pivot);
  //    vec lane  -- h_filtered = h_pred + h_node.   Information form.
  double a[N];
  const bool init_per_seq = (p.flags & SVAE_LDS_FLAG_INIT_PER_SEQ) != 0;
  const double* iJ = p.init_J + (init_per_seq ? (size_t)b * NN : 0);
  const double* ih = p.init_h + (init_per_seq ? (size_t)b * N : 0);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (mat) {
      a[i] = -2.0 * iJ[i * N + c];
      if (!INHOMOG && T > 1) a[i] -= 2.0 * p.J11[i * N + c];
    } else {
      a[i] = vec ? ih[i] + nring[NP + i] : 0.0;
    }
  }

  // =========================== forward in time: information filter ===========================
  for (int t = 0; t < T; ++t) {
    const bool last = (t == T - 1);
    if (INHOMOG && !last) {
      const size_t off = (size_t)t * p.pair_stride_t + (size_t)b * p.pair_stride_b;
      if (mat) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double j11 = -2.0 * p.J11[off + i * N + c];
          a[i] += j11;
          tJ12[i * NP + c] = -p.J12[off + i * N + c];
          tJ22[i * NP + c] = -2.0 * p.J22[off + i * N + c];
        }
      }
      __syncwarp();
    }
    const double* nd = nring + (t & 1) * NODE;

    // right-hand sides: r = J12 column (mat), q = identity column (mat -> M^-1 column)
    double r[N], q[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      r[i] = last ? 0.0 : J12s[i * NP + cm];
      q[i] = (i == c) ? 1.0 : 0.0;
    }

    // ---- square-root-free LDL' of M_t fused with the forward substitution of all right-hand
    //      sides: lanes = rows; the pivot column is broadcast through Lr; the vec lane's `a`
    //      rides along as a bordered row, which IS its forward substitution ----------------------
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (mat) Lr[k * NP + c] = a[k];
      __syncwarp();
      double row[NP];
#pragma unroll
      for (int q2 = k / 2; q2 < NP / 2; ++q2) {
        const double2 v = ld2(&Lr[k * NP + 2 * q2]);
        row[2 * q2] = v.x;
        row[2 * q2 + 1] = v.y;
      }
      const double piv = fma(-2.0, nd[k], row[k]);     // + node precision on the diagonal
      bad |= !(piv > 0.0);
      const double ik = fast_rcp(piv);
      const double ta = a[k] * ik, tr = r[k] * ik, tq = q[k] * ik;
      if (mat) Lt[k * NP + c] = ta;                      // unit-lower factor, column k
      if (vec) lacc = fma(0.5 * a[k], ta, lacc);         // 1/2 z_k^2 / d_k
#pragma unroll
      for (int j = k + 1; j < N; ++j) {
        a[j] = fma(-ta, row[j], a[j]);
        r[j] = fma(-tr, row[j], r[j]);
        q[j] = fma(-tq, row[j], q[j]);
      }
      a[k] = ta;
      r[k] = tr;
      q[k] = tq;
    }
    if (mat) {
      const double dc = fma(-2.0, nd[c], Lr[c * NP + c]);
      dprod *= dc;
      const double rs = rsqrt(dc);
      sq[c] = rs;
      if (save) mbuf[c] = dc * rs;   // sqrt(d_c), consumed below (mbuf is free until m_t is published)
    }
    if ((t & 7) == 7 || last) {
      lacc -= 0.5 * log(dprod);
      dprod = 1.0;
    }
    if (vec) {
#pragma unroll
      for (int i = 0; i < N; ++i) r[i] = a[i];
    }
    __syncwarp();
    if (smp) {
      const double* ep = nd + (2 + s) * NP;
#pragma unroll
      for (int k = 0; k < N; ++k) r[k] = ep[k] * sq[k];   // D^-1/2 eps
    }
    if (save) {
      // K = L^-1 column c: (L~^-1 e_c)_i / sqrt(d_i) = q_i sqrt(d_i)   (q is D^-1 L~^-1 e_c here)
      if (mat && valid) {
        double* kt = stash + (size_t)t * RS + (2 * N + 1 + S) * NP;
#pragma unroll
        for (int i = 0; i < N; ++i) kt[i * NP + c] = q[i] * mbuf[i];
      }
      __syncwarp();   // mbuf is rewritten with m_t below
    }
    // ---- backward substitution with the unit-upper factor L~' (in place) ----------------------
#pragma unroll
    for (int k = N - 2; k >= 0; --k) {
      double accr = 0.0, accq = 0.0;
#pragma unroll
      for (int q2 = (k + 1) / 2; q2 < NP / 2; ++q2) {
        const double2 v = ld2(&Lt[k * NP + 2 * q2]);
        if (2 * q2 > k && 2 * q2 < N) {
          accr = fma(v.x, r[2 * q2], accr);
          accq = fma(v.x, q[2 * q2], accq);
        }
        if (2 * q2 + 1 > k && 2 * q2 + 1 < N) {
          accr = fma(v.y, r[2 * q2 + 1], accr);
          accq = fma(v.y, q[2 * q2 + 1], accq);
        }
      }
      r[k] -= accr;
      q[k] -= accq;
    }
    // mat: r = M^-1 J12[:,c] = -G[:,c], q = M^-1[:,c];  vec: r = m;  smp: r = L^-T eps
    if (vec) {
#pragma unroll
      for (int i = 0; i < N; ++i) mbuf[i] = r[i];
    }
    __syncwarp();
    if (smp) {
#pragma unroll
      for (int i = 0; i < N; ++i) r[i] += mbuf[i];   // b_t = m_t + L^-T eps_t
    }

    if (!last || save) {
      // ---- stash record for the backward sweep (and, when saving, for the VJP kernel) ---------
      double* st = stash + (size_t)t * RS;
      if (valid) {
        if (mat) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            st[i * NP + c] = -r[i];
            st[N * NP + i * NP + c] = q[i];
          }
          st[2 * N * NP + c] = mbuf[c];
        } else if (smp) {
#pragma unroll
          for (int i = 0; i < N; ++i) st[(2 * N + 1 + s) * NP + i] = r[i];
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
        const double* pinit = mat ? pm + c : ndn + NP;
        const int pstr = mat ? NP : 1;
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = pinit[i * pstr];
        if (INHOMOG) {
#pragma unroll
          for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int q2 = 0; q2 < NP / 2; ++q2) {
              const double2 v = ld2(&J12s[k * NP + 2 * q2]);
              a[2 * q2] = fma(-v.x, r[k], a[2 * q2]);
              if (2 * q2 + 1 < N) a[2 * q2 + 1] = fma(-v.y, r[k], a[2 * q2 + 1]);
            }
          }
        } else {
          // information-form J12 = -(natural J12):  a -= J12_info' r  ==  a += J12_nat' r
          const double* cj12 = c_J12[p.const_slot];
#pragma unroll
          for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] = fma(cj12[k * N + i], r[k], a[i]);
          }
        }
      }
    } else {
      // terminal step: the smoothed marginal at T-1 is the filtered one
#pragma unroll
      for (int i = 0; i < N; ++i) a[i] = mat ? q[i] : r[i];
    }
    __syncwarp();
  }
  // backward-sweep state: column of [Sigma | mu | x]  (reuses the registers of a)
  double (&z)[N] = a;

  // lognorm: reduce the per-lane partials over the tile, add the constant terms
  {
    double tot = lacc;
#pragma unroll
    for (int o = LPS / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o, LPS);
    if (c == 0 && valid && p.lognorm) {
      double konst = p.init_logZ ? p.init_logZ[init_per_seq ? b : 0] : 0.0;
      if (p.pair_logZ) {
        if (!INHOMOG) {
          konst += (double)(T - 1) * p.pair_logZ[0];
        } else {
          for (int t = 0; t < T - 1; ++t)
            konst += p.pair_logZ[((size_t)t * p.pair_stride_t + (size_t)b * p.pair_stride_b) / NN];
        }
      }
      if (p.node_logZ)
        for (int t = 0; t < T; ++t) konst += p.node_logZ[(size_t)b * T + t];
      p.lognorm[b] = tot + konst;
    }
  }
  if (bad && p.status) atomicMax(p.status, (int)b + 1);

  // =========================== backward in time: smoother + sampler ==========================
  // stash records were written through the generic proxy; order them before the async-proxy reads
  __threadfence_block();
  fence_proxy_async();
  __syncwarp();
  auto fetch = [&](int t, int j) {   // j-th backward step (t = T-2-j) -> ring slot j&1
#if SVAE_LDS_TMA
    if (c == 0) {
      mbar_arrive_expect_tx(&bars[j & 1], (uint32_t)(SS * sizeof(double)));
      tma_load_1d(ring + (j & 1) * SS, stash + (size_t)t * RS, (uint32_t)(SS * sizeof(double)), &bars[j & 1]);
    }
#else
    double* dst = ring + (j & 1) * SS;
    const double* src = stash + (size_t)t * RS;
    for (int e = c; e < SS; e += LPS) dst[e] = src[e];
#endif
  };
  if (T >= 2) fetch(T - 2, 0);
  if (T >= 3) fetch(T - 3, 1);

  double acc_mid[N], acc_xxn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc_mid[i] = acc_xxn[i] = 0.0;
  double mun_c = 0.0;

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
      // C = G z  for every column of [Sigma_{t+1} | mu_{t+1} | x_{t+1}]
      {
        double cc[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 g = ld2(&G[i * NP + 2 * q2]);
            acc = fma(g.x, z[2 * q2], acc);
            if (2 * q2 + 1 < N) acc = fma(g.y, z[2 * q2 + 1], acc);
          }
          cc[i] = acc;
        }
        if (mat) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            Cs[c * NP + i] = cc[i];
            acc_xxn[i] += cc[i];     // Cov(x_t, x_{t+1})[i][c]; the mean outer product is added below
          }
        } else if (vec || smp) {
          const double* base = rec + (2 * N + (vec ? 0 : 1 + s)) * NP;
#pragma unroll
          for (int i = 0; i < N; ++i) z[i] = base[i] + cc[i];
        }
      }
      __syncwarp();
      if (mat) {
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = Minv[i * NP + c];
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const double gk = G[c * NP + k];
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 v = ld2(&Cs[k * NP + 2 * q2]);
            z[2 * q2] = fma(v.x, gk, z[2 * q2]);
            if (2 * q2 + 1 < N) z[2 * q2 + 1] = fma(v.y, gk, z[2 * q2 + 1]);   // Sigma_t[:, c]
          }
        }
      }
    }
    // ---- emit statistics for time t -----------------------------------------------------------
    if (vec) {
#pragma unroll
      for (int i = 0; i < N; ++i) mu_s[i] = z[i];
    } else if (smp) {
#pragma unroll
      for (int i = 0; i < N; ++i) xs[s * NP + i] = z[i];
    }
    __syncwarp();
    if (mat && save && valid) {
      double* sg = stash + (size_t)t * RS + (3 * N + 1 + S) * NP;
#pragma unroll
      for (int i = 0; i < N; ++i) sg[i * NP + c] = z[i];   // Sigma_t[:, c]
    }
    if (mat) {
      const double mu_c = mu_s[c];
      const size_t bt = (size_t)b * T + t;
      const bool first = (t == 0), lastt = (t == T - 1);
      // hot path: running sums only
      double diag = 0.0;
      if (!first && !lastt) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double mi = mu_s[i];
          const double exx = fma(mi, mu_c, z[i]);           // E[x_t x_t'][i][c]
          if (i == c) diag = exx;
          acc_xxn[i] = fma(mi, mun_c, acc_xxn[i]);
          acc_mid[i] += exx;
        }
      } else {
        // first / last time step: E_init, the end corrections of the pair sums
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double mi = mu_s[i];
          const double exx = fma(mi, mu_c, z[i]);
          if (i == c) diag = exx;
          if (!lastt) acc_xxn[i] = fma(mi, mun_c, acc_xxn[i]);
          if (valid) {
            const size_t o = (size_t)b * NN + i * N + c;
            if (lastt && p.E_pair_xnxn) p.E_pair_xnxn[o] = (T > 1) ? exx : 0.0;
            if (first) {
              if (p.E_init_xx) p.E_init_xx[o] = exx;
              if (p.E_pair_xx) p.E_pair_xx[o] = (T > 1) ? acc_mid[i] + exx : 0.0;
              if (p.E_pair_xxn) p.E_pair_xxn[o] = acc_xxn[i];
              if (p.E_pair_xnxn && T > 1) p.E_pair_xnxn[o] = (lastt ? 0.0 : p.E_pair_xnxn[o]) + acc_mid[i];
            }
          }
        }
        if (first && valid && p.E_init_x) p.E_init_x[(size_t)b * N + c] = mu_c;
      }
      if (valid) {
        if (p.E_node_x) p.E_node_x[bt * N + c] = mu_c;
        if (p.E_node_xx) p.E_node_xx[bt * N + c] = diag;
        if (p.samples)
          for (int ss = 0; ss < S; ++ss) p.samples[(bt * S + ss) * N + c] = xs[ss * NP + c];
        if (p.Exx_t) {   // per-time second moments (inhomogeneous E_pair): cold path
#pragma unroll
          for (int i = 0; i < N; ++i) {
            p.Exx_t[bt * NN + i * N + c] = fma(mu_s[i], mu_c, z[i]);
            if (!lastt && p.Exxn_t)
              p.Exxn_t[((size_t)b * (T - 1) + t) * NN + i * N + c] = fma(mu_s[i], mun_c, Cs[c * NP + i]);
          }
        }
      }
      mun_c = mu_c;
    }
    __syncwarp();
    if (t >= 2 && t < T - 1) fetch(t - 2, T - t);   // the slot just consumed is free: prefetch
  }
}

}  // namespace svae
