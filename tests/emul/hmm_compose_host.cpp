// Host build of svae_b200/csrc/hmm_compose_core.h (TEST INFRASTRUCTURE): the phase code of svae_hmm_compose_chunks
// and svae_hmm_chain_messages with the lanes looped.
#include <vector>
#include <algorithm>
#include "hmm_compose_core.h"

extern "C" void hmm_compose_chunks_host(int B, int T, int K, int chunk_len, const double* log_A, const double* node, double* P,
                                        double* logscale) {
  const int C = (T - 1 + chunk_len - 1) / chunk_len;
  std::vector<double> mem(svae::hmm_compose_tile_doubles(K));
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const long long tile = (long long)b * C + c;
      const int s = c * chunk_len, L = std::min(chunk_len, T - 1 - s);
      svae::hmm_compose_chunk(mem.data(), K, s, L, log_A, node + (size_t)b * T * K, P + tile * K * K, logscale + tile);
    }
}

extern "C" void hmm_chain_messages_host(int B, int C, int K, const double* log_pi, const double* P, const double* logscale,
                                        const double* node_last, double* log_alpha, double* log_beta, double* logZ) {
  std::vector<double> mem(svae::hmm_chain_tile_doubles(K));
  for (long long b = 0; b < B; ++b)
    svae::hmm_chain_messages(mem.data(), K, C, log_pi, P + b * C * K * K, logscale + b * C, node_last + b * K,
                             log_alpha + b * (C + 1) * K, log_beta + b * (C + 1) * K, logZ + b);
}

extern "C" void hmm_compose_chunks_maxplus_host(int B, int T, int K, int chunk_len, const double* log_A, const double* node,
                                                double* V) {
  const int C = (T - 1 + chunk_len - 1) / chunk_len;
  std::vector<double> mem(svae::hmm_maxplus_tile_doubles(K));
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const long long tile = (long long)b * C + c;
      const int s = c * chunk_len, L = std::min(chunk_len, T - 1 - s);
      svae::hmm_compose_chunk_maxplus(mem.data(), K, s, L, log_A, node + (size_t)b * T * K, V + tile * K * K);
    }
}

extern "C" void hmm_chain_viterbi_host(int B, int C, int K, const double* log_pi, const double* V, const double* node_last,
                                       double* delta, int* zb) {
  std::vector<double> mem(svae::hmm_chain_viterbi_tile_doubles(K));
  for (long long b = 0; b < B; ++b)
    svae::hmm_chain_viterbi(mem.data(), K, C, log_pi, V + b * C * K * K, node_last + b * K, delta + b * (C + 1) * K,
                            zb + b * (C + 1));
}
