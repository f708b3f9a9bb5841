// hmm_compose.cu -- svae_hmm_compose_chunks / svae_hmm_chain_messages: stages 1 and 2 of the parallel-in-time HMM
// E-step (hmm_compose_core.h); stage 3 is an ordinary svae_hmm_estep launch over all chunks.
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include "svae_b200.h"
#include "hmm_compose_core.h"

extern std::atomic<uint64_t> g_svae_launches;

namespace {
constexpr int kWarps = 2;

__global__ void __launch_bounds__(32 * svae::kHmmWarps) hmm_compose_kernel(int B, int T, int K, int Lc, int C,
                                                                           const double* log_A, const double* node, double* P,
                                                                           double* logscale) {
  extern __shared__ double smem[];
  const long long tile = blockIdx.x;      // one CTA per (sequence, chunk)
  const int b = (int)(tile / C), c = (int)(tile % C);
  const int s = c * Lc, L = min(Lc, T - 1 - s);
  svae::hmm_compose_chunk(smem, K, s, L, log_A, node + (size_t)b * T * K, P + tile * K * K, logscale + tile);
}

__global__ void __launch_bounds__(32 * kWarps) hmm_chain_kernel(int B, int K, int C, const double* log_pi, const double* P,
                                                                const double* logscale, const double* node_last,
                                                                double* log_alpha, double* log_beta, double* logZ) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * kWarps + warp;
  if (b >= B) return;
  svae::hmm_chain_messages(smem + (size_t)warp * svae::hmm_chain_tile_doubles(K), K, C, log_pi, P + b * C * K * K,
                           logscale + b * C, node_last + b * K, log_alpha + b * (C + 1) * K, log_beta + b * (C + 1) * K,
                           logZ + b);
}
__global__ void __launch_bounds__(32 * svae::kHmmWarps) hmm_compose_maxplus_kernel(int B, int T, int K, int Lc, int C,
                                                                                   const double* log_A, const double* node,
                                                                                   double* V) {
  extern __shared__ double smem[];
  const long long tile = blockIdx.x;      // one CTA per (sequence, chunk)
  const int b = (int)(tile / C), c = (int)(tile % C);
  const int s = c * Lc, L = min(Lc, T - 1 - s);
  svae::hmm_compose_chunk_maxplus(smem, K, s, L, log_A, node + (size_t)b * T * K, V + tile * K * K);
}

__global__ void __launch_bounds__(32 * kWarps) hmm_chain_viterbi_kernel(int B, int K, int C, const double* log_pi,
                                                                        const double* V, const double* node_last,
                                                                        double* delta, int* zb) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * kWarps + warp;
  if (b >= B) return;
  svae::hmm_chain_viterbi(smem + (size_t)warp * svae::hmm_chain_viterbi_tile_doubles(K), K, C, log_pi, V + b * C * K * K,
                          node_last + b * K, delta + b * (C + 1) * K, zb + b * (C + 1));
}
}  // namespace

extern "C" int svae_hmm_compose_chunks_maxplus(int B, int T, int K, int chunk_len, const double* log_A, const double* node,
                                               double* chunk_V, void* stream) {
  if (B <= 0 || T < 2 || K <= 0 || chunk_len < 1 || !log_A || !node || !chunk_V) return SVAE_ERR_BAD_ARG;
  if (K > SVAE_HMM_MAX_K) return SVAE_ERR_UNSUPPORTED;
  const int C = svae_lds_compose_num_chunks(T, chunk_len);
  const size_t smem = sizeof(double) * (size_t)svae::hmm_maxplus_tile_doubles(K);
  hmm_compose_maxplus_kernel<<<(unsigned)((long long)B * C), 32 * svae::kHmmWarps, smem, static_cast<cudaStream_t>(stream)>>>(
      B, T, K, chunk_len, C, log_A, node, chunk_V);
  g_svae_launches.fetch_add(1);
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}

extern "C" int svae_hmm_chain_viterbi(int B, int C, int K, const double* log_pi, const double* chunk_V,
                                      const double* node_last, double* delta_scratch, int* boundary_states, void* stream) {
  if (B <= 0 || C <= 0 || K <= 0 || !log_pi || !chunk_V || !node_last || !delta_scratch || !boundary_states)
    return SVAE_ERR_BAD_ARG;
  if (K > SVAE_HMM_MAX_K) return SVAE_ERR_UNSUPPORTED;
  const size_t smem = sizeof(double) * (size_t)kWarps * svae::hmm_chain_viterbi_tile_doubles(K);
  hmm_chain_viterbi_kernel<<<(unsigned)((B + kWarps - 1) / kWarps), 32 * kWarps, smem, static_cast<cudaStream_t>(stream)>>>(
      B, K, C, log_pi, chunk_V, node_last, delta_scratch, boundary_states);
  g_svae_launches.fetch_add(1);
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}

extern "C" int svae_hmm_compose_chunks(int B, int T, int K, int chunk_len, const double* log_A, const double* node,
                                       double* chunk_P, double* chunk_logscale, void* stream) {
  if (B <= 0 || T < 2 || K <= 0 || chunk_len < 1 || !log_A || !node || !chunk_P || !chunk_logscale) return SVAE_ERR_BAD_ARG;
  if (K > SVAE_HMM_MAX_K) return SVAE_ERR_UNSUPPORTED;
  const int C = svae_lds_compose_num_chunks(T, chunk_len);
  const size_t smem = sizeof(double) * (size_t)svae::hmm_compose_tile_doubles(K);
  const long long tiles = (long long)B * C;
  hmm_compose_kernel<<<(unsigned)tiles, 32 * svae::kHmmWarps, smem, static_cast<cudaStream_t>(stream)>>>(
      B, T, K, chunk_len, C, log_A, node, chunk_P, chunk_logscale);
  g_svae_launches.fetch_add(1);
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}

extern "C" int svae_hmm_chain_messages(int B, int C, int K, const double* log_pi, const double* chunk_P,
                                       const double* chunk_logscale, const double* node_last, double* log_alpha,
                                       double* log_beta, double* logZ, void* stream) {
  if (B <= 0 || C <= 0 || K <= 0 || !log_pi || !chunk_P || !chunk_logscale || !node_last || !log_alpha || !log_beta || !logZ)
    return SVAE_ERR_BAD_ARG;
  if (K > SVAE_HMM_MAX_K) return SVAE_ERR_UNSUPPORTED;
  const size_t smem = sizeof(double) * (size_t)kWarps * svae::hmm_chain_tile_doubles(K);
  hmm_chain_kernel<<<(unsigned)((B + kWarps - 1) / kWarps), 32 * kWarps, smem, static_cast<cudaStream_t>(stream)>>>(
      B, K, C, log_pi, chunk_P, chunk_logscale, node_last, log_alpha, log_beta, logZ);
  g_svae_launches.fetch_add(1);
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}
