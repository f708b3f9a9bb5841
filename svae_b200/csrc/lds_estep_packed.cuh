// lds_estep_packed.cuh -- the hot variant of the fused LDS E-step (homogeneous dynamics, diagonal node
// potentials, S <= 2 sample paths, no message ports): same mathematics and same stash record as
// lds_estep_fwd_kernel (lds_estep.cuh), different lane mapping.
//
// What it replaces in the reference is what lds_estep.cuh replaces:
//   svae/lds/cython_lds_inference.pyx:28-90, 149-210, 310-355 and the per-step kernels in
//   svae/lds/cython_gaussian_grads.pxd:38-82,125-144,167-188,208-296,431-508.
//
// Why a second mapping.  The general kernel gives every sequence a power-of-two tile of 16 or 32 lanes and
// spends one lane on the mean vector and one per sample path: at n = 10, S = 1 only 12 of 16 lanes of a tile do
// matrix work, the vector lanes run their own divergent instruction streams, and n = 16 needs a whole warp per
// sequence.  ncu (profiles/r01_final_fwd_summary.txt, per-phase breakdown in profiles/r02_fwd_phase_breakdown.txt)
// shows a flat profile bounded by issue slots, shared-memory instructions and dependent-issue latency together --
// the lever is instructions per sequence.  Here a sequence owns exactly N lanes (lane c = column c of every
// n x n matrix), 32/N sequences share a warp (3 at n = 10, 4 at n = 8, 2 at n = 16), and the vectors are
// DISTRIBUTED over the matrix lanes (lane c holds element c), which turns every vector recursion into a short
// product against operands the lane already holds:
//   m_t      = M^-1 h_f            ->  m_c  = sum_i Minv[i][c] h_f[i]          (q = column c of M^-1, in registers)
//   h_pred'  = -J12' M^-1 h_f      ->  hp_c = -sum_i Y[i][c] h_f[i]            (r = column c of Y = M^-1 J12)
//   L^-T eps (reference's map from noise to sample, pxd:431-454)
//            = K' eps, K = L^-1    ->  xi_c = sum_i q~_c[i] sqrt(d_i) eps_i    (q~ = D^-1 L~^-1 e_c: q before the
//                                                                               back substitution)
//   mu_t     = m_t + G_t mu_{t+1}  ->  mu_c = m_c + sum_k G[c][k] mu_{t+1}[k]  (row c of G: the same operands the
//   x_t      = b_t + G_t x_{t+1}                                                Sigma_t update loads anyway)
//   1/2 h_f' M^-1 h_f              ->  per-lane partial h_f[c] m_c, reduced once per sequence
// No lane diverges from the others, the forward substitution of h rides nowhere (it is not needed any more), and
// broadcast operands are read by all lanes of a sequence with one 16-byte shared-memory load.
#pragma once
#include "lds_estep.cuh"
#ifdef SVAE_PACKED_CLOCKS
#include <cstdio>
#endif

namespace svae {

__host__ __device__ constexpr int lds_packed_spw(int N) { return 32 / N; }

// -DSVAE_PACKED_NOG=1: the smoother gain G_t = M_t^-1 J12 is NOT written to the stash record; the backward sweep rebuilds
// what it needs of it from M_t^-1 (which it fetches anyway) and the constant-bank J12:  C = M^-1 (J12 Sigma_{t+1}),
// row c of G = J12' M^-1[:,c].  +2 n^2 multiply-adds per lane-step on constant operands, no extra shared-memory traffic,
// the record's fetched part shrinks from (2n+1+S) to (n+1+S) rows (homogeneous, non-saving launches only: the VJP kernel
// reads G from the saved record).
#ifndef SVAE_PACKED_NOG
#define SVAE_PACKED_NOG 0
#endif

__host__ __device__ constexpr bool lds_packed_nog(bool save, bool inhomog) { return (SVAE_PACKED_NOG != 0) && !save && !inhomog; }

// shared-memory doubles per sequence tile; padded so that consecutive tiles start one ROW (NP doubles) apart modulo
// 128 bytes: the rows the 32/N sequences of a warp touch with one instruction then sit side by side in bank space
// (2 wavefronts for 3 x 80 bytes at n = 10 instead of the 4 that ncu counted with a 48-byte skew)
__host__ __device__ constexpr int lds_packed_tile_doubles(int N, int S, bool inhomog = false, bool nog = false) {
  const int NP = lds_np(N);
  const int fa = 2 * N * NP + 2 * (2 + S) * NP + (inhomog ? 6 * N * NP : 0);   // Lr, Lt, node ring, 2 x {J11|J12|J22} of a step
  const int fb = 2 * ((nog ? N : 2 * N) + 1 + S) * NP + N * NP;   // stash ring (fetched part of two records), Cs
  const int raw = (3 + 2 * S) * NP /*hf, mu, sd, w, x*/ + (fa > fb ? fa : fb);
  return raw + ((NP % 16 - raw % 16) + 16) % 16;
}

// register budget: one warp per CTA; resident warps per SM chosen so that the metric batch is one wave
#ifndef SVAE_PACKED_MINB
#define SVAE_PACKED_MINB(N) ((N) <= 8 ? 16 : (N) <= 10 ? 12 : (N) <= 12 ? 10 : 8)
#endif

// dev instrumentation (-DSVAE_PACKED_CLOCKS): per-phase clock64 sums of block 0, printed at kernel end
#ifdef SVAE_PACKED_CLOCKS
#define SVAE_CLK(i) { const long long now_ = clock64(); clk_[i] += now_ - tprev_; tprev_ = now_; }
#else
#define SVAE_CLK(i)
#endif

// reciprocal square root: hardware seed + two Newton steps (pivots are positive, O(1) numbers; the library routine
// spends ~20 instructions on special cases that cannot occur past the sign check on the pivot)
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}

// INHOMOG: pair parameters per (sequence, time) (the SLDS mixture of K modes, svae/models/slds_svae.py:92-103): a step's three
// n x n blocks are staged by cp.async one step ahead (every lane copies its own column; the J12 rows the prediction
// broadcasts become visible at the end-of-step barrier), the initial state may be per sequence, the per-time second moments
// E[x_t x_t'], E[x_t x_{t+1}'] can be written out (the HMM node potentials are contracted from them).
template <int N, int S, bool SAVE, bool INHOMOG>
__global__ void __launch_bounds__(32, SVAE_PACKED_MINB(N)) lds_estep_packed_kernel(const LdsFwdParams p) {
  constexpr int NP = lds_np(N);
  constexpr int NN = N * N;
  constexpr int SPW = lds_packed_spw(N);
  constexpr int ACT = SPW * N;                                   // lanes that own a column
  constexpr unsigned MASK = ACT == 32 ? 0xffffffffu : ((1u << ACT) - 1u);
  constexpr int SS = (2 * N + 1 + S) * NP;                       // fetched prefix of a stash record
  constexpr int NODE = (2 + S) * NP;
  constexpr bool NOG = lds_packed_nog(SAVE, INHOMOG);
  constexpr int FOFF = NOG ? N * NP : 0;                         // fetched part of a record: doubles [FOFF, SS)
  constexpr int SLOT = SS - FOFF;                                // ring slot
  constexpr int TILE = lds_packed_tile_doubles(N, S, INHOMOG, NOG);
  constexpr int SA = S > 0 ? S : 1;                              // array extent (no zero-length arrays)
  constexpr int PB = 3 * N * NP;                                 // one staged step of pair parameters: J11 | J12 | J22
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x;
  const int T = p.T;
#ifdef SVAE_PACKED_CLOCKS
  long long clk_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev_ = clock64();
#endif

  // ---- CTA-shared homogeneous pair parameters, information form ------------------------------
  double* shJ11 = smem;            // J22 + J11: what every non-terminal prediction lands in
  double* shJ12 = shJ11 + N * NP;
  double* shJ22 = shJ12 + N * NP;
  if (T > 1 && !INHOMOG) {
    for (int e = lane; e < NN; e += 32) {
      const int i = e / N, j = e % N;
      shJ11[i * NP + j] = -2.0 * p.J22[e] - 2.0 * p.J11[e];
      shJ12[i * NP + j] = -p.J12[e];
      shJ22[i * NP + j] = -2.0 * p.J22[e];
    }
  }
  __syncwarp();
  if (lane >= ACT) return;         // (32 mod N) spare lanes

  const int sub = lane / N, c = lane - sub * N;
  const long long seq = (long long)blockIdx.x * SPW + sub;
  const long long b = seq < p.B ? seq : (long long)p.B - 1;
  const bool valid = seq < p.B && (!p.active || p.active[b] != 0);
  constexpr bool save = SAVE;
  constexpr int RS = (SAVE ? 4 * N + 1 + S : 2 * N + 1 + S) * NP;   // == lds_record_stride(N, S, SAVE)

  uint64_t* bars = reinterpret_cast<uint64_t*>(shJ22 + N * NP);   // 2 warp-level mbarriers (stash ring slots)
  double* tile0 = shJ22 + N * NP + 2;
  double* tile = tile0 + (size_t)sub * TILE;
  double* hfv = tile;              // h_filtered of this step (broadcast operand)
  double* mu_s = hfv + NP;         // mu_{t+1} during the backward sweep
  double* sdv = mu_s + NP;         // sqrt(d_k)
  double* wv = sdv + NP;           // sqrt(d_k) eps_k per sample path
  double* xs = wv + S * NP;        // x_{t+1} per sample path
  double* uni = xs + S * NP;       // ---- phase-shared ----
  double* Lr = uni;                // Lr[k*NP+j]: pivot row k of the LDL' elimination (j >= k), Lr[k*NP+k] = d_k
  double* Lt = Lr + N * NP;        // Lt[i*NP+k] = L~[i][k], unit lower factor, row-major
  double* nring = Lt + N * NP;     // 2 x { J_node | h_node | eps rows }
  double* pst = nring + 2 * NODE;  // INHOMOG: 2 x { J11 | J12 | J22 } natural form, row-major [i][j]
  double* ring = uni;              // backward sweep: 2 stash records
  double* Cs = ring + 2 * SLOT;    // Cs[k*NP+i] = C[i][k]

  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncwarp(MASK);

  double* stash = p.stash + (size_t)b * T * RS;
  const double* nJ = p.node_J + (size_t)b * T * N + c;
  const double* nh = p.node_h + (size_t)b * T * N + c;
  const double* eps = S > 0 ? p.eps + (size_t)b * T * S * N + c : nullptr;
  // per-time outputs, element c of this sequence (the packed launcher requires every output buffer)
  const bool stats = p.E_node_x != nullptr;          // statistic buffers are optional (natural_lds_sample: samples only)
  double* oEx = stats ? p.E_node_x + (size_t)b * T * N + c : nullptr;
  double* oExx = stats ? p.E_node_xx + (size_t)b * T * N + c : nullptr;
  double* oSmp = S > 0 ? p.samples + (size_t)b * T * S * N + c : nullptr;

  // node potentials (+ noise) of time t -> ring slot t&1; every lane copies and later reads its own element only
  auto stage_node = [&](int t) {
    double* dst = nring + (t & 1) * NODE;
    cp_async8(dst + c, nJ + t * N);
    cp_async8(dst + NP + c, nh + t * N);
#pragma unroll
    for (int s = 0; s < S; ++s) cp_async8(dst + (2 + s) * NP + c, eps + (t * S + s) * N);
    cp_async_commit();
  };

  const bool init_per_seq = INHOMOG && (p.flags & SVAE_LDS_FLAG_INIT_PER_SEQ) != 0;
  auto stage_pair = [&](int t) {              // pair parameters of step t -> pst slot t&1 (own column of each block)
    if (INHOMOG) {
      const size_t off = (size_t)t * p.pair_stride_t + (size_t)b * p.pair_stride_b + c;
      double* dst = pst + (t & 1) * PB + c;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        cp_async8(dst + i * NP, p.J11 + off + i * N);
        cp_async8(dst + N * NP + i * NP, p.J12 + off + i * N);
        cp_async8(dst + 2 * N * NP + i * NP, p.J22 + off + i * N);
      }
      cp_async_commit();
    }
  };

  double lacc = 0.0;     // - 1/2 sum log d (this lane's pivots)
  double lq = 0.0;       // sum_t h_f[c] m_t[c]
  double dprod = 1.0;
  int badbits = 0;

  if (INHOMOG && T > 1) stage_pair(0);
  stage_node(0);
  if (T > 1) stage_node(1); else cp_async_commit();
  cp_async_wait<1>();                          // everything but node 1 has landed (own elements only are read here)

  // a: column c of M_t = J_pred + J11 (the node precision joins at the pivot), information form
  double a[N];
  const double* iJ = p.init_J + (init_per_seq ? (size_t)b * NN : 0);
  double hpred = p.init_h[(init_per_seq ? (size_t)b * N : 0) + c];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a[i] = -2.0 * iJ[i * N + c];
    if (T > 1) a[i] -= 2.0 * (INHOMOG ? pst[i * NP + c] : p.J11[i * N + c]);
  }
  double mu = 0.0, x[SA];
#pragma unroll
  for (int s = 0; s < SA; ++s) x[s] = 0.0;

  // =========================== forward in time: information filter ===========================
  for (int t = 0; t < T; ++t) {
    SVAE_CLK(0)
    const bool last = (t == T - 1);
    const double* nd = nring + (t & 1) * NODE;
    const double ndJ = nd[c];
    const double hf = hpred + nd[NP + c];
    double ec[SA];
#pragma unroll
    for (int s = 0; s < S; ++s) ec[s] = nd[(2 + s) * NP + c];
    hfv[c] = hf;                    // visible to the other lanes after the elimination's barriers

    // right-hand sides: r = column c of J12 (-> Y = M^-1 J12), q = e_c (-> column c of M^-1)
    const double* pb = pst + (t & 1) * PB;      // INHOMOG: this step's J11 | J12 | J22
    if (INHOMOG && t + 1 < T - 1) stage_pair(t + 1);   // the other slot is free since the end of the previous step
    double r[N], q[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      r[i] = last ? 0.0 : (INHOMOG ? -pb[N * NP + i * NP + c] : shJ12[i * NP + c]);
      q[i] = (i == c) ? 1.0 : 0.0;
    }

    // ---- square-root-free LDL' of M_t fused with the forward substitution of both right-hand sides -------
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double v = a[k];
      if (k == c) v = fma(-2.0, ndJ, v);            // + node precision on the diagonal
      Lr[k * NP + c] = v;
      __syncwarp(MASK);
      double row[NP];
#pragma unroll
      for (int q2 = k / 2; q2 < NP / 2; ++q2) {
        const double2 w = ld2(&Lr[k * NP + 2 * q2]);
        row[2 * q2] = w.x;
        row[2 * q2 + 1] = w.y;
      }
      badbits |= __double2hiint(row[k]);
      const double ik = fast_rcp(row[k]);
      const double ta = v * ik, tr = r[k] * ik, tq = q[k] * ik;
#pragma unroll
      for (int j = k + 1; j < N; ++j) {
        a[j] = fma(-ta, row[j], a[j]);
        r[j] = fma(-tr, row[j], r[j]);
        q[j] = fma(-tq, row[j], q[j]);
      }
      a[k] = ta;                     // L~[c][k] (k < c; the rest of the row is never read)
      r[k] = tr;
      q[k] = tq;
    }
    // row c of the unit lower factor, one 16-byte store per pair (a column store per pivot costs 6 wavefronts each)
#pragma unroll
    for (int q2 = 0; q2 < NP / 2; ++q2) st2(&Lt[c * NP + 2 * q2], a[2 * q2], (2 * q2 + 1 < N) ? a[2 * q2 + 1] : 0.0);
    SVAE_CLK(1)
    {
      const double dc = Lr[c * NP + c];             // this lane's pivot (its own store)
      dprod *= dc;
      if ((t & 7) == 7 || last) {     // pivots are O(1): eight of them cannot leave the double range
        lacc -= 0.5 * log(dprod);
        dprod = 1.0;
      }
      if (S > 0 || save) {
        const double sd = dc * fast_rsqrt(dc);
        if (save) sdv[c] = sd;
#pragma unroll
        for (int s = 0; s < S; ++s) wv[s * NP + c] = sd * ec[s];
      }
    }
    __syncwarp(MASK);
    // xi = L^-T eps = sum_i q~[i] sqrt(d_i) eps_i  (q~ = D^-1 L~^-1 e_c, i.e. q BEFORE the back substitution)
    double xi[SA];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      double x0 = 0.0, x1 = 0.0;
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        const double2 w = ld2(&wv[s * NP + 2 * q2]);
        x0 = fma(q[2 * q2], w.x, x0);
        if (2 * q2 + 1 < N) x1 = fma(q[2 * q2 + 1], w.y, x1);
      }
      xi[s] = x0 + x1;
    }
    if (save) {
      // K = L^-1, column c: (L~^-1 e_c)_i / sqrt(d_i) = q~_i sqrt(d_i); consumed by the VJP kernel
      double* kt = stash + (size_t)t * RS + (2 * N + 1 + S) * NP;
      if (valid) {
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 w = ld2(&sdv[2 * q2]);
          kt[(2 * q2) * NP + c] = q[2 * q2] * w.x;
          if (2 * q2 + 1 < N) kt[(2 * q2 + 1) * NP + c] = q[2 * q2 + 1] * w.y;
        }
      }
    }
    SVAE_CLK(2)
    // ---- back substitution with the unit upper factor L~', column-oriented (independent accumulators) ----
#pragma unroll
    for (int i = N - 1; i >= 1; --i) {
#pragma unroll
      for (int q2 = 0; q2 < (i + 1) / 2; ++q2) {
        const double2 w = ld2(&Lt[i * NP + 2 * q2]);
        r[2 * q2] = fma(-w.x, r[i], r[2 * q2]);
        q[2 * q2] = fma(-w.x, q[i], q[2 * q2]);
        if (2 * q2 + 1 < i) {
          r[2 * q2 + 1] = fma(-w.y, r[i], r[2 * q2 + 1]);
          q[2 * q2 + 1] = fma(-w.y, q[i], q[2 * q2 + 1]);
        }
      }
    }
    SVAE_CLK(3)
    // r = Y[:,c] = -G[:,c],  q = M^-1[:,c].   m_c = <q, h_f>,  (J12' m)_c = <r, h_f>
    double m, hq;
    {
      double m0 = 0.0, m1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        const double2 w = ld2(&hfv[2 * q2]);
        m0 = fma(q[2 * q2], w.x, m0);
        h0 = fma(r[2 * q2], w.x, h0);
        if (2 * q2 + 1 < N) {
          m1 = fma(q[2 * q2 + 1], w.y, m1);
          h1 = fma(r[2 * q2 + 1], w.y, h1);
        }
      }
      m = m0 + m1;
      hq = h0 + h1;
    }
    lq = fma(hf, m, lq);
    double bs[SA];
#pragma unroll
    for (int s = 0; s < S; ++s) bs[s] = m + xi[s];   // b_t = m_t + L^-T eps_t

    if ((!last || save) && valid) {
      // stash record: G' (row c = column c of G, contiguous) | M^-1 (symmetric: row c) | m | b
      double* st = stash + (size_t)t * RS;
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        if (!NOG) st2(&st[c * NP + 2 * q2], -r[2 * q2], (2 * q2 + 1 < N) ? -r[2 * q2 + 1] : 0.0);
        st2(&st[N * NP + c * NP + 2 * q2], q[2 * q2], (2 * q2 + 1 < N) ? q[2 * q2 + 1] : 0.0);
      }
      st[2 * N * NP + c] = m;
#pragma unroll
      for (int s = 0; s < S; ++s) st[(2 * N + 1 + s) * NP + c] = bs[s];
    }
    SVAE_CLK(4)
    if (!last) {
      cp_async_wait<0>();            // own elements of step t+1 (node potentials, pair parameters) have landed
      if (t + 2 < T) stage_node(t + 2);
      // predict straight into next step's M:  J22 + J11 - J12' Y   (information-form J12 = -natural J12)
      if (INHOMOG) {
        const double* pn = pst + ((t + 1) & 1) * PB;              // next step's J11 (own column: no barrier needed)
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = -2.0 * pb[2 * N * NP + i * NP + c] - ((t + 1 < T - 1) ? 2.0 * pn[i * NP + c] : 0.0);
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 w = ld2(&pb[N * NP + k * NP + 2 * q2]);   // natural J12[k][:]: visible since the last barrier
            a[2 * q2] = fma(w.x, r[k], a[2 * q2]);
            if (2 * q2 + 1 < N) a[2 * q2 + 1] = fma(w.y, r[k], a[2 * q2 + 1]);
          }
        }
      } else {
        const double* pm = (t + 1 == T - 1) ? shJ22 : shJ11;
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = pm[i * NP + c];
        const double* cj12 = c_J12[p.const_slot];
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int i = 0; i < N; ++i) a[i] = fma(cj12[k * N + i], r[k], a[i]);
        }
      }
      hpred = -hq;
    } else {
      // terminal step: the smoothed marginal at T-1 is the filtered one
#pragma unroll
      for (int i = 0; i < N; ++i) a[i] = q[i];
      mu = m;
#pragma unroll
      for (int s = 0; s < S; ++s) x[s] = bs[s];
    }
    __syncwarp(MASK);                // hfv / wv / Lr / Lt are rewritten by the next step
    SVAE_CLK(5)
  }
  double (&z)[N] = a;                // backward-sweep state: column c of Sigma_t

  // ---- lognorm: reduce the per-lane partials over the sequence's lanes, add the constant terms ------------
  {
    const double part = fma(0.5, lq, lacc);
    hfv[c] = part;
    __syncwarp(MASK);
    if (c == 0 && valid) {
      double tot = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) tot += hfv[i];
      double konst = p.init_logZ ? p.init_logZ[init_per_seq ? b : 0] : 0.0;
      if (p.pair_logZ) {
        if (!INHOMOG) {
          konst += (double)(T - 1) * p.pair_logZ[0];
        } else {
          const size_t lz0 = (size_t)b * p.pair_stride_b / NN, lzs = (size_t)p.pair_stride_t / NN;   // strides are multiples of n^2
          for (int t = 0; t < T - 1; ++t) konst += p.pair_logZ[lz0 + (size_t)t * lzs];
        }
      }
      if (p.node_logZ)
        for (int t = 0; t < T; ++t) konst += p.node_logZ[(size_t)b * T + t];
      if (p.lognorm) p.lognorm[b] = tot + konst;
    }
    // negative pivot (sign bit) or a NaN that reached the normaliser
    if ((badbits < 0 || !(part == part)) && p.status) atomicMax(p.status, (int)b + 1);
  }

  // =========================== backward in time: smoother + sampler ==========================
  __threadfence_block();             // stash records were written through the generic proxy
  fence_proxy_async();
  __syncwarp(MASK);
  // One lane issues the bulk copies of ALL the warp's sequences onto one warp-level mbarrier: every address below is
  // a function of blockIdx and compile-time constants, so the copies run on the uniform datapath back to back (a
  // copy per sequence issued by that sequence's own lane costs an ELECT / R2UR loop over the lanes instead).
  constexpr int ring_off = (3 + 2 * S) * NP;   // == ring - tile
  auto fetch = [&](int t, int j) {   // j-th backward step (t = T-2-j) -> ring slot j&1
#if SVAE_LDS_TMA
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[j & 1], (uint32_t)(SPW * SLOT * sizeof(double)));
#pragma unroll
      for (int s2 = 0; s2 < SPW; ++s2) {
        long long bs = (long long)blockIdx.x * SPW + s2;
        if (bs >= p.B) bs = (long long)p.B - 1;
        tma_load_1d(tile0 + s2 * TILE + ring_off + (j & 1) * SLOT, p.stash + ((size_t)bs * T + t) * RS + FOFF,
                    (uint32_t)(SLOT * sizeof(double)), &bars[j & 1]);
      }
    }
#else
    double* dst = ring + (j & 1) * SLOT;
    const double* src = stash + (size_t)t * RS + FOFF;
    for (int e = c; e < SLOT; e += N) dst[e] = src[e];
#endif
  };
  if (T >= 2) fetch(T - 2, 0);
  if (T >= 3) fetch(T - 3, 1);

  double acc_mid[N], acc_xxn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc_mid[i] = acc_xxn[i] = 0.0;
  double mun = 0.0;                  // mu_{t+1}[c]

  for (int t = T - 1; t >= 0; --t) {
    SVAE_CLK(6)
    if (t < T - 1) {
      const int j = T - 2 - t;
      const double* rec = ring + (j & 1) * SLOT - FOFF;     // record-relative offsets below; [0, FOFF) is not fetched
#if SVAE_LDS_TMA
      mbar_wait(&bars[j & 1], (uint32_t)((j >> 1) & 1));
#else
      __syncwarp(MASK);
#endif
      SVAE_CLK(7)
      const double* G = rec;                    // G[k*NP+i] = G_t[i][k]
      const double* Minv = rec + N * NP;
      // C[:,c] = G Sigma_{t+1}[:,c]
      double cc[N], gk[N];
#pragma unroll
      for (int i = 0; i < N; ++i) cc[i] = 0.0;
      if constexpr (NOG) {
        // G = M^-1 J12 (natural-form J12 in the constant bank):  C[:,c] = M^-1 (J12 Sigma_{t+1}[:,c])
        const double* cj12 = c_J12[p.const_slot];
        double w[N];
#pragma unroll
        for (int j = 0; j < N; ++j) w[j] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int j = 0; j < N; ++j) w[j] = fma(cj12[j * N + k], z[k], w[j]);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 g = ld2(&Minv[j * NP + 2 * q2]);          // row j of the symmetric M^-1
            cc[2 * q2] = fma(g.x, w[j], cc[2 * q2]);
            if (2 * q2 + 1 < N) cc[2 * q2 + 1] = fma(g.y, w[j], cc[2 * q2 + 1]);
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
          for (int q2 = 0; q2 < NP / 2; ++q2) {
            const double2 g = ld2(&G[k * NP + 2 * q2]);
            cc[2 * q2] = fma(g.x, z[k], cc[2 * q2]);
            if (2 * q2 + 1 < N) cc[2 * q2 + 1] = fma(g.y, z[k], cc[2 * q2 + 1]);
          }
        }
        // row c of G
#pragma unroll
        for (int k = 0; k < N; ++k) gk[k] = G[k * NP + c];
      }
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        st2(&Cs[c * NP + 2 * q2], cc[2 * q2], (2 * q2 + 1 < N) ? cc[2 * q2 + 1] : 0.0);
        acc_xxn[2 * q2] += cc[2 * q2];            // Cov(x_t, x_{t+1})[i][c]; the mean outer product joins below
        if (2 * q2 + 1 < N) acc_xxn[2 * q2 + 1] += cc[2 * q2 + 1];
      }
      if constexpr (NOG) {
        const double* cj12 = c_J12[p.const_slot];
        // row c of G:  G[c][k] = sum_j M^-1[j][c] J12[j][k]
#pragma unroll
        for (int k = 0; k < N; ++k) gk[k] = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const double zj = Minv[j * NP + c];
#pragma unroll
          for (int k = 0; k < N; ++k) gk[k] = fma(cj12[j * N + k], zj, gk[k]);
        }
      }
      // element c of  mu_t = m_t + G mu_{t+1},  x_t = b_t + G x_{t+1}  (row c of G: the operands the Sigma update uses anyway)
      double mnew = rec[2 * N * NP + c], xnew[SA];
#pragma unroll
      for (int s = 0; s < S; ++s) xnew[s] = rec[(2 * N + 1 + s) * NP + c];
      {
        double e0 = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 w = ld2(&mu_s[2 * q2]);
          mnew = fma(gk[2 * q2], w.x, mnew);
          if (2 * q2 + 1 < N) e0 = fma(gk[2 * q2 + 1], w.y, e0);
        }
        mnew += e0;
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        double e0 = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 w = ld2(&xs[s * NP + 2 * q2]);
          xnew[s] = fma(gk[2 * q2], w.x, xnew[s]);
          if (2 * q2 + 1 < N) e0 = fma(gk[2 * q2 + 1], w.y, e0);
        }
        xnew[s] += e0;
      }
      __syncwarp(MASK);              // Cs complete; every lane is done with mu_{t+1}, x_{t+1}
      SVAE_CLK(8)
      // Sigma_t[:,c] = M^-1[:,c] + C G'[:,c]
#pragma unroll
      for (int i = 0; i < N; ++i) z[i] = Minv[i * NP + c];
#pragma unroll
      for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int q2 = 0; q2 < NP / 2; ++q2) {
          const double2 w = ld2(&Cs[k * NP + 2 * q2]);
          z[2 * q2] = fma(w.x, gk[k], z[2 * q2]);
          if (2 * q2 + 1 < N) z[2 * q2 + 1] = fma(w.y, gk[k], z[2 * q2 + 1]);
        }
      }
      mun = mu;
      mu = mnew;
#pragma unroll
      for (int s = 0; s < S; ++s) x[s] = xnew[s];
    }
    SVAE_CLK(9)
    mu_s[c] = mu;
#pragma unroll
    for (int s = 0; s < S; ++s) xs[s * NP + c] = x[s];
    __syncwarp(MASK);                // also: every lane is done with ring slot j&1

    // ---- emit statistics for time t -----------------------------------------------------------
    const bool first = (t == 0), lastt = (t == T - 1);
    if (save && valid) {
      double* sg = stash + (size_t)t * RS + (3 * N + 1 + S) * NP;
#pragma unroll
      for (int i = 0; i < N; ++i) sg[i * NP + c] = z[i];   // Sigma_t[:, c]
    }
    double diag = 0.0;
    if (!first && !lastt) {
      // hot path: running sums only
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        const double2 w = ld2(&mu_s[2 * q2]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 2 * q2 + h;
          if (i < N) {
            const double mi = h ? w.y : w.x;
            const double exx = fma(mi, mu, z[i]);           // E[x_t x_t'][i][c]
            if (i == c) diag = exx;
            acc_xxn[i] = fma(mi, mun, acc_xxn[i]);
            acc_mid[i] += exx;
          }
        }
      }
    } else {
      // first / last time step: E_init, the end corrections of the pair sums
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double mi = mu_s[i];
        const double exx = fma(mi, mu, z[i]);
        if (i == c) diag = exx;
        if (!lastt) acc_xxn[i] = fma(mi, mun, acc_xxn[i]);
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
      if (first && valid && p.E_init_x) p.E_init_x[(size_t)b * N + c] = mu;
    }
    if (INHOMOG && p.Exx_t && valid) {
      // per-time second moments: E[x_t x_t'][:,c] = Sigma_t[:,c] + mu_t mu_t[c],  E[x_t x_{t+1}'][:,c] = C[:,c] + mu_t mu_{t+1}[c]
      double* ox = p.Exx_t + ((size_t)b * T + t) * NN + c;
      double* oxn = (p.Exxn_t && !lastt) ? p.Exxn_t + ((size_t)b * (T - 1) + t) * NN + c : nullptr;
#pragma unroll
      for (int q2 = 0; q2 < NP / 2; ++q2) {
        const double2 w = ld2(&mu_s[2 * q2]);
        const double2 cv = ld2(&Cs[c * NP + 2 * q2]);               // C[:,c]: this lane's own row of the tile
        ox[(2 * q2) * N] = fma(w.x, mu, z[2 * q2]);
        if (oxn) oxn[(2 * q2) * N] = fma(w.x, mun, cv.x);
        if (2 * q2 + 1 < N) {
          ox[(2 * q2 + 1) * N] = fma(w.y, mu, z[2 * q2 + 1]);
          if (oxn) oxn[(2 * q2 + 1) * N] = fma(w.y, mun, cv.y);
        }
      }
    }
    if (valid) {
      if (stats) {
        oEx[t * N] = mu;
        oExx[t * N] = diag;
      }
#pragma unroll
      for (int s = 0; s < S; ++s) oSmp[(t * S + s) * N] = x[s];
    }
    if (t >= 2 && t < T - 1) fetch(t - 2, T - t);   // the slot just consumed is free: prefetch
    SVAE_CLK(10)
  }
#ifdef SVAE_PACKED_CLOCKS
  if (blockIdx.x == 0 && lane == 0)
    printf("[clk N=%d S=%d B=%d T=%d] per step: top %lld | ldl %lld | dc/xi %lld | backsub %lld | m/stash %lld | predict %lld || "
           "loop %lld | mbar %lld | Cgz+vec %lld | sigma %lld | emit %lld\n", N, S, p.B, T, clk_[0] / T, clk_[1] / T, clk_[2] / T,
           clk_[3] / T, clk_[4] / T, clk_[5] / T, clk_[6] / T, clk_[7] / T, clk_[8] / T, clk_[9] / T, clk_[10] / T);
#endif
}

}  // namespace svae
