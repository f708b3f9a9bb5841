// api.cu -- the extern "C" boundary of libsvae_b200.so (see include/svae_b200.h).
#include <atomic>
#include <cstdlib>
#include "svae_b200.h"
#include "launch.h"

std::atomic<uint64_t> g_svae_launches{0};

extern "C" const char* svae_b200_version(void) { return "svae_b200 0.1 sm_100a"; }
extern "C" uint64_t svae_b200_launch_count(void) { return g_svae_launches.load(); }

namespace {
typedef cudaError_t (*lds_fwd_fn)(const svae::LdsFwdParams&, int, int, bool, cudaStream_t);
const lds_fwd_fn kLdsFwd[SVAE_LDS_MAX_N + 1] = {
    nullptr,
    svae::lds_fwd_launch_n1, svae::lds_fwd_launch_n2, svae::lds_fwd_launch_n3, svae::lds_fwd_launch_n4,
    svae::lds_fwd_launch_n5, svae::lds_fwd_launch_n6, svae::lds_fwd_launch_n7, svae::lds_fwd_launch_n8,
    svae::lds_fwd_launch_n9, svae::lds_fwd_launch_n10, svae::lds_fwd_launch_n11, svae::lds_fwd_launch_n12,
    svae::lds_fwd_launch_n13, svae::lds_fwd_launch_n14, svae::lds_fwd_launch_n15, svae::lds_fwd_launch_n16};

typedef cudaError_t (*lds_bwd_fn)(const svae::LdsBwdParams&, int, cudaStream_t);
const lds_bwd_fn kLdsBwd[SVAE_LDS_MAX_N + 1] = {
    nullptr,
    svae::lds_bwd_launch_n1, svae::lds_bwd_launch_n2, svae::lds_bwd_launch_n3, svae::lds_bwd_launch_n4,
    svae::lds_bwd_launch_n5, svae::lds_bwd_launch_n6, svae::lds_bwd_launch_n7, svae::lds_bwd_launch_n8,
    svae::lds_bwd_launch_n9, svae::lds_bwd_launch_n10, svae::lds_bwd_launch_n11, svae::lds_bwd_launch_n12,
    svae::lds_bwd_launch_n13, svae::lds_bwd_launch_n14, svae::lds_bwd_launch_n15, svae::lds_bwd_launch_n16};

int pick_lps(int n, int S) {   // one column per lane (VJP kernel; forward kernel with SVAE_LDS_CPL=1)
  if (n + 1 + S <= 16) return 16;
  if (n + 1 + S <= 32) return 32;
  return 0;
}
// forward kernel: two column slots per lane, the smallest tile that holds n + 1 + S slots
int pick_lps2(int n, int S) {
  for (int lps = 4; lps <= 32; lps *= 2)
    if (2 * lps >= n + 1 + S) return lps;
  return 0;
}
int forward_cpl() {
#ifdef SVAE_LDS_ENABLE_CPL2
  static int cpl = [] { const char* e = getenv("SVAE_LDS_CPL"); return (e && atoi(e) == 2) ? 2 : 1; }();
  return cpl;
#else
  return 1;
#endif
}
}  // namespace

extern "C" size_t svae_lds_estep_workspace_bytes(int B, int T, int n, int S, int flags) {
  if (B <= 0 || T <= 0 || n <= 0 || S < 0) return 0;
  const bool save = (flags & SVAE_LDS_SAVE_FOR_BWD) != 0;
  return sizeof(double) * (size_t)B * (size_t)T * (size_t)svae::lds_record_stride(n, S, save);
}

extern "C" int svae_lds_estep_fwd_masked(
    int B, int T, int n, int S, int flags,
    const double* init_J, const double* init_h, const double* init_logZ,
    const double* J11, const double* J12, const double* J22, const double* pair_logZ,
    long long pair_stride_t, long long pair_stride_b,
    const double* node_J, const double* node_h, const double* node_logZ,
    const double* eps,
    double* E_init_xx, double* E_init_x,
    double* E_pair_xx, double* E_pair_xxn, double* E_pair_xnxn,
    double* E_node_xx, double* E_node_x,
    double* samples, double* lognorm,
    double* Exx_t, double* Exxn_t,
    double* msg_J_pred, double* msg_h_pred, const double* in_J_filt, const double* in_h_filt,
    const double* node_J_dense, const double* term_J, const int* active,
    int* status, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || T <= 0 || n <= 0 || S < 0) return SVAE_ERR_BAD_ARG;
  if (!init_J || !init_h || !node_J || !node_h || !workspace) return SVAE_ERR_BAD_ARG;
  if (T > 1 && (!J11 || !J12 || !J22)) return SVAE_ERR_BAD_ARG;
  if (S > 0 && !eps) return SVAE_ERR_BAD_ARG;
  if (n > SVAE_LDS_MAX_N) return SVAE_ERR_UNSUPPORTED;
  const int cpl = forward_cpl();
  const int lps = cpl == 2 ? pick_lps2(n, S) : pick_lps(n, S);
  if (!lps) return SVAE_ERR_UNSUPPORTED;
  if (workspace_bytes < svae_lds_estep_workspace_bytes(B, T, n, S, flags)) return SVAE_ERR_WORKSPACE;

  svae::LdsFwdParams p;
  p.B = B; p.T = T; p.S = S; p.flags = flags; p.const_slot = 0;
  p.init_J = init_J; p.init_h = init_h; p.init_logZ = init_logZ;
  p.J11 = J11; p.J12 = J12; p.J22 = J22; p.pair_logZ = pair_logZ;
  p.pair_stride_t = pair_stride_t; p.pair_stride_b = pair_stride_b;
  p.node_J = node_J; p.node_h = node_h; p.node_logZ = node_logZ;
  p.eps = eps;
  p.E_init_xx = E_init_xx; p.E_init_x = E_init_x;
  p.E_pair_xx = E_pair_xx; p.E_pair_xxn = E_pair_xxn; p.E_pair_xnxn = E_pair_xnxn;
  p.E_node_xx = E_node_xx; p.E_node_x = E_node_x;
  p.samples = samples; p.lognorm = lognorm;
  const bool per_t = (flags & SVAE_LDS_PER_T_PAIR_STATS) != 0;
  p.Exx_t = per_t ? Exx_t : nullptr; p.Exxn_t = per_t ? Exxn_t : nullptr;
  p.msg_J_pred = msg_J_pred; p.msg_h_pred = msg_h_pred; p.in_J_filt = in_J_filt; p.in_h_filt = in_h_filt;
  p.node_J_dense = node_J_dense; p.term_J = term_J; p.active = active;
  p.status = status;
  p.stash = static_cast<double*>(workspace);
  const bool inhomog = (pair_stride_t != 0 || pair_stride_b != 0) && T > 1;
  cudaError_t e = kLdsFwd[n](p, cpl, lps, inhomog, static_cast<cudaStream_t>(stream));
  g_svae_launches.fetch_add(1);
  return e == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}

extern "C" int svae_lds_estep_fwd_ex(
    int B, int T, int n, int S, int flags,
    const double* init_J, const double* init_h, const double* init_logZ,
    const double* J11, const double* J12, const double* J22, const double* pair_logZ,
    long long pair_stride_t, long long pair_stride_b,
    const double* node_J, const double* node_h, const double* node_logZ,
    const double* eps,
    double* E_init_xx, double* E_init_x,
    double* E_pair_xx, double* E_pair_xxn, double* E_pair_xnxn,
    double* E_node_xx, double* E_node_x,
    double* samples, double* lognorm,
    double* Exx_t, double* Exxn_t,
    double* msg_J_pred, double* msg_h_pred, const double* in_J_filt, const double* in_h_filt,
    const double* node_J_dense, const double* term_J,
    int* status, void* workspace, size_t workspace_bytes, void* stream) {
  return svae_lds_estep_fwd_masked(B, T, n, S, flags, init_J, init_h, init_logZ, J11, J12, J22, pair_logZ, pair_stride_t,
                                   pair_stride_b, node_J, node_h, node_logZ, eps, E_init_xx, E_init_x, E_pair_xx, E_pair_xxn,
                                   E_pair_xnxn, E_node_xx, E_node_x, samples, lognorm, Exx_t, Exxn_t, msg_J_pred, msg_h_pred,
                                   in_J_filt, in_h_filt, node_J_dense, term_J, nullptr, status, workspace, workspace_bytes,
                                   stream);
}

extern "C" int svae_lds_estep_fwd(
    int B, int T, int n, int S, int flags,
    const double* init_J, const double* init_h, const double* init_logZ,
    const double* J11, const double* J12, const double* J22, const double* pair_logZ,
    long long pair_stride_t, long long pair_stride_b,
    const double* node_J, const double* node_h, const double* node_logZ,
    const double* eps,
    double* E_init_xx, double* E_init_x,
    double* E_pair_xx, double* E_pair_xxn, double* E_pair_xnxn,
    double* E_node_xx, double* E_node_x,
    double* samples, double* lognorm,
    double* Exx_t, double* Exxn_t,
    int* status, void* workspace, size_t workspace_bytes, void* stream) {
  return svae_lds_estep_fwd_ex(B, T, n, S, flags, init_J, init_h, init_logZ, J11, J12, J22, pair_logZ, pair_stride_t,
                               pair_stride_b, node_J, node_h, node_logZ, eps, E_init_xx, E_init_x, E_pair_xx, E_pair_xxn,
                               E_pair_xnxn, E_node_xx, E_node_x, samples, lognorm, Exx_t, Exxn_t, nullptr, nullptr, nullptr,
                               nullptr, nullptr, nullptr, status, workspace, workspace_bytes, stream);
}

extern "C" int svae_lds_estep_bwd_ports(
    int B, int T, int n, int S, int flags,
    const double* E_node_x, const double* samples, const double* eps,
    const double* g_samples,
    const double* g_E_init_xx, const double* g_E_init_x,
    const double* g_E_pair_xx, const double* g_E_pair_xxn, const double* g_E_pair_xnxn,
    const double* g_Exx_t, const double* g_Exxn_t,
    const double* g_E_node_xx, const double* g_E_node_x, const double* g_lognorm,
    double* g_node_J, double* g_node_h, double* g_node_J_dense,
    double* g_msg_M, double* g_msg_h,
    const double* in_g_M_pred, const double* in_g_h_pred, const double* in_g_M_filt, const double* in_g_h_filt,
    void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || T <= 0 || n <= 0 || S < 0) return SVAE_ERR_BAD_ARG;
  const bool s1 = (flags & SVAE_LDS_BWD_MESSAGES_OUT) != 0, s2 = (flags & SVAE_LDS_BWD_MESSAGES_IN) != 0;
  if (s1 && s2) return SVAE_ERR_BAD_ARG;
  if (s1 && (!g_msg_M || !g_msg_h)) return SVAE_ERR_BAD_ARG;
  if (!E_node_x || !workspace) return SVAE_ERR_BAD_ARG;
  if (!s1 && ((!g_node_J && !g_node_J_dense) || !g_node_h)) return SVAE_ERR_BAD_ARG;
  if (S > 0 && (!samples || !eps)) return SVAE_ERR_BAD_ARG;
  if (n > SVAE_LDS_MAX_N) return SVAE_ERR_UNSUPPORTED;
  const int lps = pick_lps(n, S);
  if (!lps) return SVAE_ERR_UNSUPPORTED;
  if (workspace_bytes < svae_lds_estep_workspace_bytes(B, T, n, S, flags | SVAE_LDS_SAVE_FOR_BWD)) return SVAE_ERR_WORKSPACE;
  svae::LdsBwdParams p;
  p.B = B; p.T = T; p.S = S; p.flags = flags;
  p.E_node_x = E_node_x; p.samples = samples; p.eps = eps;
  p.g_samples = g_samples;
  p.g_E_init_xx = g_E_init_xx; p.g_E_init_x = g_E_init_x;
  p.g_E_pair_xx = g_E_pair_xx; p.g_E_pair_xxn = g_E_pair_xxn; p.g_E_pair_xnxn = g_E_pair_xnxn;
  p.g_Exx_t = g_Exx_t; p.g_Exxn_t = g_Exxn_t;
  p.g_E_node_xx = g_E_node_xx; p.g_E_node_x = g_E_node_x; p.g_lognorm = g_lognorm;
  p.g_node_J = g_node_J; p.g_node_h = g_node_h; p.g_node_J_dense = g_node_J_dense;
  p.g_msg_M = g_msg_M; p.g_msg_h = g_msg_M ? g_msg_h : nullptr;
  p.in_g_M_pred = in_g_M_pred; p.in_g_h_pred = in_g_h_pred; p.in_g_M_filt = in_g_M_filt; p.in_g_h_filt = in_g_h_filt;
  p.flags = (flags & ~(SVAE_LDS_BWD_MESSAGES_OUT | SVAE_LDS_BWD_MESSAGES_IN)) | (s1 ? SVAE_LDS_BWD_SWEEP1_ONLY : 0) |
            (s2 ? SVAE_LDS_BWD_SWEEP2_ONLY : 0);
  p.stash = static_cast<double*>(workspace);
  cudaError_t e = kLdsBwd[n](p, lps, static_cast<cudaStream_t>(stream));
  g_svae_launches.fetch_add(1);
  return e == cudaSuccess ? SVAE_OK : SVAE_ERR_CUDA;
}

extern "C" int svae_lds_estep_bwd_ex(
    int B, int T, int n, int S, int flags,
    const double* E_node_x, const double* samples, const double* eps,
    const double* g_samples,
    const double* g_E_init_xx, const double* g_E_init_x,
    const double* g_E_pair_xx, const double* g_E_pair_xxn, const double* g_E_pair_xnxn,
    const double* g_Exx_t, const double* g_Exxn_t,
    const double* g_E_node_xx, const double* g_E_node_x, const double* g_lognorm,
    double* g_node_J, double* g_node_h, double* g_node_J_dense,
    void* workspace, size_t workspace_bytes, void* stream) {
  return svae_lds_estep_bwd_ports(B, T, n, S, flags & ~(SVAE_LDS_BWD_MESSAGES_OUT | SVAE_LDS_BWD_MESSAGES_IN), E_node_x, samples,
                                  eps, g_samples, g_E_init_xx, g_E_init_x, g_E_pair_xx, g_E_pair_xxn, g_E_pair_xnxn, g_Exx_t,
                                  g_Exxn_t, g_E_node_xx, g_E_node_x, g_lognorm, g_node_J, g_node_h, g_node_J_dense, nullptr,
                                  nullptr, nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int svae_lds_estep_bwd(
    int B, int T, int n, int S, int flags,
    const double* E_node_x, const double* samples, const double* eps,
    const double* g_samples,
    const double* g_E_init_xx, const double* g_E_init_x,
    const double* g_E_pair_xx, const double* g_E_pair_xxn, const double* g_E_pair_xnxn,
    const double* g_Exx_t, const double* g_Exxn_t,
    const double* g_E_node_xx, const double* g_E_node_x, const double* g_lognorm,
    double* g_node_J, double* g_node_h,
    void* workspace, size_t workspace_bytes, void* stream) {
  if (!g_node_J) return SVAE_ERR_BAD_ARG;
  return svae_lds_estep_bwd_ex(B, T, n, S, flags, E_node_x, samples, eps, g_samples, g_E_init_xx, g_E_init_x, g_E_pair_xx,
                               g_E_pair_xxn, g_E_pair_xnxn, g_Exx_t, g_Exxn_t, g_E_node_xx, g_E_node_x, g_lognorm, g_node_J,
                               g_node_h, nullptr, workspace, workspace_bytes, stream);
}
