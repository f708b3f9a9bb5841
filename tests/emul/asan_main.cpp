// AddressSanitizer / UBSan run over the host builds of every phase-coded kernel (TEST INFRASTRUCTURE): the tile memory
// is a heap vector of exactly the size the CUDA launchers request as shared memory, so an out-of-tile index on the
// device shows up here as a heap-buffer-overflow.  Built and run by tests/test_emul_asan.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "lds_compose_core.h"
#include "hmm_compose_core.h"
#include "hmm_wide_core.h"
#include <cstdint>

static double rnd() { return (double)rand() / RAND_MAX - 0.5; }

int main() {
  srand(1);
  // ---- LDS compose, every latent dimension, ragged chunks, homogeneous and per-step pair parameters -------------
  for (int n = 1; n <= 16; ++n) {
    const int T = 23, Lc = 5, C = (T - 1 + Lc - 1) / Lc;
    for (int inhomog = 0; inhomog < 2; ++inhomog) {
      const int reps = inhomog ? T - 1 : 1;
      std::vector<double> J11(reps * n * n), J12(reps * n * n), J22(reps * n * n), lz(reps, 0.1), nJ(T * n), nh(T * n), nz(T, 0.0);
      for (int r = 0; r < reps; ++r)
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            J11[(r * n + i) * n + j] = (i == j) ? -0.5 * 0.3 : 0.0;
            J12[(r * n + i) * n + j] = (i == j) ? 0.4 : 0.01 * rnd();
            J22[(r * n + i) * n + j] = (i == j) ? -0.5 * 1.0 : 0.0;
          }
      for (auto& x : nJ) x = -0.5 * (0.5 + std::fabs(rnd()));
      for (auto& x : nh) x = rnd();
      std::vector<double> mem(svae::compose_tile_doubles(n));
      std::vector<double> o11(n * n), o12(n * n), o22(n * n), oe1(n), oe2(n);
      double oz;
      for (int c = 0; c < C; ++c) {
        const int s = c * Lc, L = std::min(Lc, T - 1 - s);
        const int bad = svae::compose_chunk_n(n, mem.data(), s, L, J11.data(), J12.data(), J22.data(), lz.data(),
                                              inhomog ? (long long)n * n : 0, nJ.data(), nh.data(), nz.data(), o11.data(),
                                              o12.data(), o22.data(), &oz, oe1.data(), oe2.data());
        if (bad || !(oz == oz)) { std::printf("LDS compose n=%d failed\n", n); return 1; }
      }
    }
  }
  // ---- HMM sum-product and max-plus stages -----------------------------------------------------------------------
  for (int K : {1, 2, 7, 17, 32}) {
    const int T = 41, Lc = 6, C = (T - 1 + Lc - 1) / Lc;
    std::vector<double> lA(K * K), lp(K), node(T * K);
    for (auto& x : lA) x = rnd();
    for (auto& x : lp) x = rnd();
    for (auto& x : node) x = 3 * rnd();
    std::vector<double> P((size_t)C * K * K), ls(C), V((size_t)C * K * K), la((C + 1) * K), lb((C + 1) * K), delta((C + 1) * K);
    std::vector<int> zb(C + 1);
    double logZ;
    {
      std::vector<double> mem(svae::hmm_compose_tile_doubles(K)), mem2(svae::hmm_maxplus_tile_doubles(K));
      for (int c = 0; c < C; ++c) {
        const int s = c * Lc, L = std::min(Lc, T - 1 - s);
        svae::hmm_compose_chunk(mem.data(), K, s, L, lA.data(), node.data(), P.data() + (size_t)c * K * K, &ls[c]);
        svae::hmm_compose_chunk_maxplus(mem2.data(), K, s, L, lA.data(), node.data(), V.data() + (size_t)c * K * K);
      }
    }
    {
      std::vector<double> mem(svae::hmm_chain_tile_doubles(K)), mem2(svae::hmm_chain_viterbi_tile_doubles(K));
      svae::hmm_chain_messages(mem.data(), K, C, lp.data(), P.data(), ls.data(), node.data() + (size_t)(T - 1) * K, la.data(),
                               lb.data(), &logZ);
      svae::hmm_chain_viterbi(mem2.data(), K, C, lp.data(), V.data(), node.data() + (size_t)(T - 1) * K, delta.data(), zb.data());
    }
    if (!(logZ == logZ)) { std::printf("HMM K=%d failed\n", K); return 1; }
    for (int z : zb)
      if (z < 0 || z >= K) { std::printf("HMM K=%d boundary state out of range\n", K); return 1; }
  }
  // ---- wide HMM (K > 32): E-step and Viterbi, exact tile / workspace / output extents -------------------------------
  for (int K : {33, 40, 64, 97}) {
    for (int T : {1, 2, 19}) {
      std::vector<double> lA(K * K), lp(K), node((size_t)T * K);
      for (auto& x : lA) x = rnd();
      for (auto& x : lp) x = rnd();
      for (auto& x : node) x = 3 * rnd();
      std::vector<double> ws(svae::hmm_wide_workspace_doubles(K)), pm(svae::hmm_wide_prep_tile_doubles(K));
      svae::hmm_wide_prep(pm.data(), K, lA.data(), ws.data(), ws.data() + (size_t)K * K, ws.data() + 2 * (size_t)K * K);
      std::vector<double> mem(svae::hmm_wide_tile_doubles(K)), E0(K), Et((size_t)K * K), Ez((size_t)T * K);
      double logZ;
      svae::hmm_wide_sequence(mem.data(), K, T, lp.data(), ws.data(), ws.data() + (size_t)K * K, ws[2 * (size_t)K * K],
                              node.data(), &logZ, E0.data(), Et.data(), Ez.data(), false);
      double tot = 0.0;
      for (double v : Et) tot += v;
      if (!(logZ == logZ) || std::fabs(tot - (T - 1)) > 1e-9 * (T + 1)) { std::printf("wide HMM K=%d T=%d failed\n", K, T); return 1; }
      std::vector<double> vm(svae::hmm_wide_viterbi_tile_doubles(K));
      std::vector<uint16_t> back((size_t)T * K);
      std::vector<int32_t> path(T);
      svae::hmm_wide_viterbi(vm.data(), K, T, lp.data(), lA.data(), node.data(), path.data(), back.data());
      for (int z : path)
        if (z < 0 || z >= K) { std::printf("wide HMM K=%d path out of range\n", K); return 1; }
    }
  }
  std::printf("asan run ok\n");
  return 0;
}
