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
//     the compile-time N; cross-lane traffic is broadcast reads from shared memory.
//   * filter step t (same arithmetic result as _natural_condition_diag + _natural_predict):
//       M_t = J_pred + diag(J_node) + J11        (information form)
//       M_t = L~ D L~'                           square-root-free LDL' (N reciprocals, no sqrt chain)
//       [Y | m] = M_t^-1 [J12 | h_f]             one augmented two-sided solve, lane = rhs column
//       J_pred' = J22 - J12' Y, h_pred' = -J12' m
//       lognorm += 1/2 h_f' m - 1/2 sum log d_k
//     and, off the recursion's critical path, everything the backward sweep needs:
//       G_t = -Y (smoother gain), M_t^-1, m_t, b_t = m_t + L^-T eps_t   -> "stash" in HBM.
//   * backward step t is then pure multiply-add on the stash (no factorisation, no solve):
//       [C | mu_t | x_t] = G_t [Sigma_{t+1} | mu_{t+1} | x_{t+1}] + [0 | m_t | b_t]
//       Sigma_t = M_t^-1 + C G_t'
//       E[x_t x_{t+1}'] = C + mu_t mu_{t+1}',  E[x_t x_t'] = Sigma_t + mu_t mu_t'
//     which equals the reference's information-form _rts_backward_step (pxd:240-296) in exact
//     arithmetic and is a sum of PSD terms numerically (no J22 - temp'temp cancellation).
//   * expected sufficient statistics are reduced over time in registers and written once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svae {

struct LdsFwdParams {
  int B, T, S, flags;
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

// doubles per (sequence, time) stash record: G (N*N) | Minv (N*N) | m (N) | b (S*N), padded to even
__host__ __device__ inline int lds_stash_stride(int N, int S) {
  int s = 2 * N * N + N + S * N;
  return (s + 1) & ~1;
}

// shared-memory doubles per sequence tile
__host__ __device__ inline int lds_tile_doubles(int N, int S, bool inhomog) {
  // Lr | Gs | Cs | hbuf | sq | mbuf | mu_s | ebuf(S*N) | xs(S*N) | vb((1+S)*N) [| J11 J12 J22]
  int d = 3 * N * N + 4 * N + 2 * S * N + (1 + S) * N + (inhomog ? 3 * N * N : 0);
  return (d + 1) & ~1;
}

template <int N, int LPS, bool INHOMOG>
__global__ void __launch_bounds__(128) lds_estep_fwd_kernel(const LdsFwdParams p) {
  constexpr int SPW = 32 / LPS;
  constexpr int NN = N * N;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPS, c = lane % LPS;
  const int S = p.S, T = p.T;
  const long long seq = ((long long)blockIdx.x * nwarps + warp) * SPW + sub;
  const bool valid = seq < p.B;
  const long long b = valid ? seq : (long long)p.B - 1;
  const bool mat = c < N, vec = c == N, smp = (c > N) && (c <= N + S);
  const int s = c - N - 1;
  const int SS = lds_stash_stride(N, S);

  // ---- shared memory carve-up -------------------------------------------------------------
  double* shJ11 = smem;           // CTA-shared homogeneous pair parameters, information form
  double* shJ12 = shJ11 + NN;
  double* shJ22 = shJ12 + NN;
  double* tile = shJ22 + NN + (size_t)(warp * SPW + sub) * lds_tile_doubles(N, S, INHOMOG);
  double* Lr = tile;              // Lr[k*N+i]: raw LDL' column k (i >= k), Lr[k*N+k] = d_k
  double* Gs = Lr + NN;           // Gs[i*N+k] = G[i][k]
  double* Cs = Gs + NN;           // Cs[k*N+i] = C[i][k]
  double* hbuf = Cs + NN;         // node h for this step
  double* sq = hbuf + N;          // 1/sqrt(d_k)
  double* mbuf = sq + N;          // m_t
  double* mu_s = mbuf + N;        // mu_t
  double* ebuf = mu_s + N;        // eps[s][k]
  double* xs = ebuf + S * N;      // x_t[s][k]
  double* vb = xs + S * N;        // staged m | b from the stash
  double* tJ11 = vb + (1 + S) * N;  // per-tile pair parameters (INHOMOG only)
  double* tJ12 = tJ11 + NN;
  double* tJ22 = tJ12 + NN;
  const double* J11s = INHOMOG ? tJ11 : shJ11;
  const double* J12s = INHOMOG ? tJ12 : shJ12;
  const double* J22s = INHOMOG ? tJ22 : shJ22;

  if (!INHOMOG) {
    for (int e = threadIdx.x; e < NN; e += blockDim.x) {
      shJ11[e] = -2.0 * p.J11[e];
      shJ12[e] = -p.J12[e];
      shJ22[e] = -2.0 * p.J22[e];
    }
  }
  __syncthreads();

  double* stash = p.stash + (size_t)b * T * SS;
  const double* nJ = p.node_J + (size_t)b * T * N;
  const double* nh = p.node_h + (size_t)b * T * N;
  const double* eps = p.eps ? p.eps + (size_t)b * T * S * N : nullptr;

  // [J_pred | h_pred] column owned by this lane (information form)
  double col[N];
#pragma unroll
  for (int i = 0; i < N; ++i) col[i] = mat ? -2.0 * p.init_J[i * N + c] : (vec ? p.init_h[i] : 0.0);

  double lacc = 0.0;   // per-lane partial of lognorm (reduced over the tile at the end)
  bool bad = false;
  double z[N];         // backward-sweep state: column of [Sigma | mu | x]

  // =========================== forward in time: information filter ===========================
  for (int t = 0; t < T; ++t) {
    const bool last = (t == T - 1);
    double jo = 0.0;
    if (mat) {
      jo = -2.0 * nJ[(size_t)t * N + c];
      hbuf[c] = nh[(size_t)t * N + c];
      for (int ss = 0; ss < S; ++ss) ebuf[ss * N + c] = eps[((size_t)t * S + ss) * N + c];
    }
    if (INHOMOG && !last) {
      const size_t off = (size_t)t * p.pair_stride_t + (size_t)b * p.pair_stride_b;
      if (mat) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          tJ11[i * N + c] = -2.0 * p.J11[off + i * N + c];
          tJ12[i * N + c] = -p.J12[off + i * N + c];
          tJ22[i * N + c] = -2.0 * p.J22[off + i * N + c];
        }
      }
    }
    __syncwarp();

    // a = column c of M_t (mat lanes); hf = h_filtered (vec lane)
    double a[N], r[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      a[i] = col[i] + (last ? 0.0 : J11s[i * N + (mat ? c : 0)]) + (i == c ? jo : 0.0);
      r[i] = vec ? col[i] + hbuf[i] : (last ? 0.0 : J12s[i * N + (mat ? c : 0)]);
    }

    // ---- square-root-free LDL' of M_t, lanes = rows, columns broadcast through Lr ----------
    double invd[N];
    double dc = 1.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (mat) Lr[k * N + c] = a[k];
      __syncwarp();
      const double piv = Lr[k * N + k];
      bad |= !(piv > 0.0);
      invd[k] = 1.0 / piv;
      if (k == c) dc = piv;
      const double tk = a[k] * invd[k];
#pragma unroll
      for (int j = k + 1; j < N; ++j) a[j] = fma(-tk, Lr[k * N + j], a[j]);
    }
    if (mat) {
      lacc -= 0.5 * log(dc);
      sq[c] = sqrt(1.0 / dc);
    }
    __syncwarp();

    // ---- augmented solve  M_t [Y | m] = [J12 | h_f]  (forward, scale, backward) -------------
    double w[N], y[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      w[k] = r[k] * invd[k];
      if (vec) lacc = fma(0.5 * r[k], w[k], lacc);
#pragma unroll
      for (int i = k + 1; i < N; ++i) r[i] = fma(-Lr[k * N + i], w[k], r[i]);
    }
    if (smp) {
#pragma unroll
      for (int k = 0; k < N; ++k) w[k] = ebuf[s * N + k] * sq[k];   // D^-1/2 eps
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
      double acc = 0.0;
#pragma unroll
      for (int i = k + 1; i < N; ++i) acc = fma(Lr[k * N + i], y[i], acc);
      y[k] = fma(-invd[k], acc, w[k]);
    }
    // mat: y = M^-1 J12[:,c] = -G[:,c];  vec: y = m;  smp: y = L^-T eps
    if (vec) {
#pragma unroll
      for (int i = 0; i < N; ++i) mbuf[i] = y[i];
    }
    __syncwarp();
    if (smp) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        y[i] += mbuf[i];   // b_t = m_t + L^-T eps_t
        xs[s * N + i] = y[i];
      }
    }

    // ---- M_t^-1 column c (second solve, rhs = e_c) -----------------------------------------
    double q[N], u[N];
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double wk = q[k] * invd[k];
#pragma unroll
      for (int i = k + 1; i < N; ++i) q[i] = fma(-Lr[k * N + i], wk, q[i]);
      q[k] = wk;
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
      double acc = 0.0;
#pragma unroll
      for (int i = k + 1; i < N; ++i) acc = fma(Lr[k * N + i], u[i], acc);
      u[k] = fma(-invd[k], acc, q[k]);
    }

    if (!last) {
      // ---- stash for the backward sweep -----------------------------------------------------
      __syncwarp();   // xs visible
      double* st = stash + (size_t)t * SS;
      if (mat && valid) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          st[i * N + c] = -y[i];
          st[NN + i * N + c] = u[i];
        }
        st[2 * NN + c] = mbuf[c];
        for (int ss = 0; ss < S; ++ss) st[2 * NN + N + ss * N + c] = xs[ss * N + c];
      }
      // ---- predict: [J_pred' | h_pred'] = [J22 | 0] - J12' [Y | m] ----------------------------
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double acc = mat ? J22s[i * N + c] : 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(-J12s[k * N + i], y[k], acc);
        col[i] = acc;
      }
    } else {
      // terminal step: the smoothed marginal at T-1 is the filtered one
#pragma unroll
      for (int i = 0; i < N; ++i) z[i] = mat ? u[i] : y[i];
    }
    __syncwarp();
  }

  // lognorm: reduce the per-lane partials over the tile, add the constant terms
  {
    double tot = lacc;
#pragma unroll
    for (int o = LPS / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o, LPS);
    if (c == 0 && valid && p.lognorm) {
      double konst = p.init_logZ ? p.init_logZ[0] : 0.0;
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
  double acc_mid[N], acc_xxn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc_mid[i] = acc_xxn[i] = 0.0;
  double mun_c = 0.0;
  double cc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) cc[i] = 0.0;

  for (int t = T - 1; t >= 0; --t) {
    if (t < T - 1) {
      const double* st = stash + (size_t)t * SS;
      double minv[N];
      if (mat) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          Gs[i * N + c] = st[i * N + c];
          minv[i] = st[NN + i * N + c];
        }
        vb[c] = st[2 * NN + c];
        for (int ss = 0; ss < S; ++ss) vb[N + ss * N + c] = st[2 * NN + N + ss * N + c];
      }
      __syncwarp();
      // C = G z  for every column of [Sigma_{t+1} | mu_{t+1} | x_{t+1}]
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(Gs[i * N + k], z[k], acc);
        cc[i] = acc;
      }
      if (mat) {
#pragma unroll
        for (int i = 0; i < N; ++i) Cs[c * N + i] = cc[i];
      } else if (vec || smp) {
        const double* base = vb + (vec ? 0 : N + s * N);
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = base[i] + cc[i];
      }
      __syncwarp();
      if (mat) {
        double grow[N];
#pragma unroll
        for (int k = 0; k < N; ++k) grow[k] = Gs[c * N + k];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double acc = minv[i];
#pragma unroll
          for (int k = 0; k < N; ++k) acc = fma(Cs[k * N + i], grow[k], acc);
          z[i] = acc;   // Sigma_t[i][c]
        }
      }
    }
    // ---- emit statistics for time t -----------------------------------------------------------
    if (vec) {
#pragma unroll
      for (int i = 0; i < N; ++i) mu_s[i] = z[i];
    } else if (smp) {
#pragma unroll
      for (int i = 0; i < N; ++i) xs[s * N + i] = z[i];
    }
    __syncwarp();
    if (mat) {
      const double mu_c = mu_s[c];
      double exx[N];
      double diag = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        exx[i] = fma(mu_s[i], mu_c, z[i]);
        if (i == c) diag = exx[i];
      }
      if (valid) {
        const size_t bt = (size_t)b * T + t;
        if (p.E_node_x) p.E_node_x[bt * N + c] = mu_c;
        if (p.E_node_xx) p.E_node_xx[bt * N + c] = diag;
        if (p.samples)
          for (int ss = 0; ss < S; ++ss) p.samples[(bt * S + ss) * N + c] = xs[ss * N + c];
        if (p.Exx_t) {
#pragma unroll
          for (int i = 0; i < N; ++i) p.Exx_t[bt * NN + i * N + c] = exx[i];
        }
        if (t < T - 1 && p.Exxn_t) {
#pragma unroll
          for (int i = 0; i < N; ++i)
            p.Exxn_t[((size_t)b * (T - 1) + t) * NN + i * N + c] = fma(mu_s[i], mun_c, cc[i]);
        }
      }
      if (t < T - 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) acc_xxn[i] += fma(mu_s[i], mun_c, cc[i]);
      }
      if (t == T - 1) {
        if (valid && p.E_pair_xnxn) {
#pragma unroll
          for (int i = 0; i < N; ++i) p.E_pair_xnxn[(size_t)b * NN + i * N + c] = (T > 1) ? exx[i] : 0.0;
        }
      } else if (t > 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) acc_mid[i] += exx[i];
      }
      if (t == 0 && valid) {
        if (p.E_init_x) p.E_init_x[(size_t)b * N + c] = mu_c;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const size_t o = (size_t)b * NN + i * N + c;
          if (p.E_init_xx) p.E_init_xx[o] = exx[i];
          if (p.E_pair_xx) p.E_pair_xx[o] = (T > 1) ? acc_mid[i] + exx[i] : 0.0;
          if (p.E_pair_xxn) p.E_pair_xxn[o] = acc_xxn[i];
          if (p.E_pair_xnxn && T > 1) p.E_pair_xnxn[o] += acc_mid[i];
        }
      }
      mun_c = mu_c;
    }
    __syncwarp();
  }
}

}  // namespace svae
