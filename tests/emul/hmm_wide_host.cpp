// Host build of svae_b200/csrc/hmm_wide_core.h (TEST INFRASTRUCTURE): the phase code of svae_hmm_estep_wide and of the
// K > 32 branch of svae_hmm_viterbi with the threads looped.
#include <vector>
#include <cstdint>
#include "hmm_wide_core.h"

extern "C" void hmm_wide_estep_host(int B, int T, int K, int fwd_only, const double* log_pi, const double* log_A,
                                    const double* node, double* logZ, double* Ez0, double* Etrans, double* Ez) {
  std::vector<double> ws(svae::hmm_wide_workspace_doubles(K));
  {
    std::vector<double> mem(svae::hmm_wide_prep_tile_doubles(K));
    svae::hmm_wide_prep(mem.data(), K, log_A, ws.data(), ws.data() + (size_t)K * K, ws.data() + 2 * (size_t)K * K);
  }
  std::vector<double> mem(svae::hmm_wide_tile_doubles(K));
  for (long long b = 0; b < B; ++b)
    svae::hmm_wide_sequence(mem.data(), K, T, log_pi, ws.data(), ws.data() + (size_t)K * K, ws[2 * (size_t)K * K],
                            node + (size_t)b * T * K, logZ + b, Ez0 ? Ez0 + (size_t)b * K : nullptr,
                            Etrans ? Etrans + (size_t)b * K * K : nullptr, Ez ? Ez + (size_t)b * T * K : nullptr, fwd_only != 0);
}

extern "C" void hmm_wide_viterbi_host(int B, int T, int K, const double* log_pi, const double* log_A, const double* node,
                                      int32_t* path) {
  std::vector<double> mem(svae::hmm_wide_viterbi_tile_doubles(K));
  std::vector<uint16_t> back((size_t)T * K);
  for (long long b = 0; b < B; ++b)
    svae::hmm_wide_viterbi(mem.data(), K, T, log_pi, log_A, node + (size_t)b * T * K, path + (size_t)b * T, back.data());
}
