// hmm_wide.cu -- svae_hmm_estep_wide and the K > 32 branch of svae_hmm_viterbi: one CTA per sequence, thread = state
// (hmm_wide_core.h).  Replaces, for more than 32 states, what hmm_estep.cu replaces for up to 32:
// svae/hmm/cython_hmm_inference.pyx:93-166 (hmm_logZ / hmm_logZ_grad), svae/hmm/hmm_inference.py:21-63.
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include "svae_b200.h"
#include "hmm_wide_core.h"

extern std::atomic<uint64_t> g_svae_launches;

namespace {

__global__ void __launch_bounds__(svae::kHmmWideMaxK) hmm_wide_prep_kernel(int K, const double* log_A, double* ws) {
  extern __shared__ double smem[];
  svae::hmm_wide_prep(smem, K, log_A, ws, ws + (size_t)K * K, ws + 2 * (size_t)K * K);
}

__global__ void __launch_bounds__(svae::kHmmWideMaxK) hmm_wide_estep_kernel(int B, int T, int K, int fwd_only, const double* log_pi, const double* ws,
                                      const double* node, const int* active, double* logZ, double* Ez0, double* Etrans,
                                      double* Ez) {
  extern __shared__ double smem[];
  const long long b = blockIdx.x;
  if (active && active[b] == 0) return;          // CTA-uniform
  const double* P = ws;
  const double* PT = ws + (size_t)K * K;
  const double gmax = ws[2 * (size_t)K * K];
  svae::hmm_wide_sequence(smem, K, T, log_pi, P, PT, gmax, node + (size_t)b * T * K, logZ ? logZ + b : nullptr,
                          Ez0 ? Ez0 + (size_t)b * K : nullptr, Etrans ? Etrans + (size_t)b * K * K : nullptr,
                          Ez ? Ez + (size_t)b * T * K : nullptr, fwd_only != 0);
}

__global__ void __launch_bounds__(svae::kHmmWideMaxK) hmm_wide_viterbi_kernel(int B, int T, int K, const double* log_pi, const double* log_A, const double* node,
                                        int32_t* path, uint16_t* back) {
  extern __shared__ double smem[];
  const long long b = blockIdx.x;
  svae::hmm_wide_viterbi(smem, K, T, log_pi, log_A, node + (size_t)b * T * K, path + (size_t)b * T, back + (size_t)b * T * K);
}

template <typename Kern>
cudaError_t allow_smem(Kern kern, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

// hmm_tc.cu: the tensor-core path for SVAE_HMM_TC_MIN_K <= K <= SVAE_HMM_MAX_K
size_t svae_hmm_tc_workspace_bytes(int B, int T, int K);
bool svae_hmm_tc_enabled();
int svae_hmm_estep_tc_launch(int B, int T, int K, bool fwd_only, const double* log_pi, const double* log_A, const double* node,
                             const int* active, double* logZ, double* Ez0, double* Etrans, double* Ez, void* workspace,
                             cudaStream_t s);

extern "C" size_t svae_hmm_wide_workspace_bytes(int B, int T, int K) {
  if (B <= 0 || T <= 0 || K < SVAE_HMM_TC_MIN_K || K > SVAE_HMM_WIDE_MAX_K) return 0;
  if (K <= SVAE_HMM_MAX_K) return svae_hmm_tc_workspace_bytes(B, T, K);
  return sizeof(double) * (size_t)svae::hmm_wide_workspace_doubles(K);
}

extern "C" int svae_hmm_estep_wide(int B, int T, int K, int flags, const double* log_pi, const double* log_A,
                                   const double* node, const int* active, double* logZ, double* Ez0, double* Etrans,
                                   double* Ez, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || T <= 0 || K <= 0) return SVAE_ERR_BAD_ARG;
  if (!log_pi || !log_A || !node) return SVAE_ERR_BAD_ARG;
  const bool fwd_only = (flags & SVAE_HMM_LOGZ_ONLY) != 0;
  if (!fwd_only && !Ez) return SVAE_ERR_BAD_ARG;
  if (K <= SVAE_HMM_MAX_K) {
    // 17..32 states with the caller's workspace: the tensor-core kernels; otherwise the register-resident kernel
    // (the log-normaliser alone leaves nothing in the workspace: no buffer needed)
    const bool ws_ok = fwd_only || (workspace && workspace_bytes >= svae_hmm_wide_workspace_bytes(B, T, K));
    if (K >= SVAE_HMM_TC_MIN_K && ws_ok && svae_hmm_tc_enabled())
      return svae_hmm_estep_tc_launch(B, T, K, fwd_only, log_pi, log_A, node, active, logZ, Ez0, Etrans, Ez, workspace,
                                      static_cast<cudaStream_t>(stream));
    return svae_hmm_estep_masked(B, T, K, flags, log_pi, log_A, node, active, logZ, Ez0, Etrans, Ez, stream);
  }
  if (K > SVAE_HMM_WIDE_MAX_K) return SVAE_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < svae_hmm_wide_workspace_bytes(B, T, K)) return SVAE_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* ws = static_cast<double*>(workspace);
  const int KP = svae::hmm_wide_kp(K);
  const size_t prep_smem = sizeof(double) * svae::hmm_wide_prep_tile_doubles(K);
  const size_t smem = sizeof(double) * svae::hmm_wide_tile_doubles(K);
  if (allow_smem(hmm_wide_estep_kernel, smem) != cudaSuccess) return SVAE_ERR_CUDA;
  hmm_wide_prep_kernel<<<1, KP, prep_smem, s>>>(K, log_A, ws);
  hmm_wide_estep_kernel<<<(unsigned)B, KP, smem, s>>>(B, T, K, fwd_only ? 1 : 0, log_pi, ws, node, active, logZ, Ez0, Etrans, Ez);
  g_svae_launches.fetch_add(2);
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}

// called by svae_hmm_viterbi (hmm_estep.cu) for K > SVAE_HMM_MAX_K
int svae_hmm_viterbi_wide_launch(int B, int T, int K, const double* log_pi, const double* log_A, const double* node,
                                 int32_t* path, void* workspace, cudaStream_t s) {
  const int KP = svae::hmm_wide_kp(K);
  const size_t smem = sizeof(double) * svae::hmm_wide_viterbi_tile_doubles(K);
  hmm_wide_viterbi_kernel<<<(unsigned)B, KP, smem, s>>>(B, T, K, log_pi, log_A, node, path, static_cast<uint16_t*>(workspace));
  g_svae_launches.fetch_add(1);
  return cudaGetLastError() == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}
