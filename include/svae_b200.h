/* svae_b200.h -- C-ABI of the B200-native SVAE local E-step engine (libsvae_b200.so).
 *
 * Drop-in boundary ("native, primitive-level"): these entry points are what the
 * reference's Cython primitives
 *     svae/lds/cython_lds_inference.pyx:28-90   natural_filter_forward_general
 *     svae/lds/cython_lds_inference.pyx:149-210 natural_smoother_general (+ _compute_stats)
 *     svae/lds/cython_lds_inference.pyx:310-355 natural_sample_backward
 *     svae/lds/cython_lds_inference.pyx:92-145, 236-306, 357-409   the three *_grad VJPs
 *     svae/hmm/cython_hmm_inference.pyx:93-166  hmm_logZ / hmm_logZ_grad
 *     svae/hmm/hmm_inference.py:21-41,54-63     hmm_estep / hmm_viterbi (pyhsmm)
 * compute one sequence at a time on the CPU; here each call processes a BATCH of B independent
 * sequences on the current CUDA device.  Plain pointers and sizes only -- no torch types.
 *
 * Conventions (identical to the reference's external arrays):
 *   - all arrays are fp64, C-order (row-major), densely packed, DEVICE pointers;
 *   - Gaussian natural parameters are (-1/2 J, h): `init_J`, `J11`, `J22`, `node_J` hold -1/2 x
 *     (information matrix), `J12` holds (Sigma^-1 A)' exactly as svae/lds/gaussian.py:130-143;
 *   - node potentials are diagonal: node_J (B,T,n), node_h (B,T,n), node_logZ (B,T) or NULL
 *     (the Cython path is diagonal-only, cython_lds_inference.pyx:43,67-70);
 *   - buffers are owned by the caller; kernels never allocate; calls are asynchronous on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream) and re-entrant per stream;
 *   - return value: 0 on success, a negative SVAE_ERR_* for host-side argument errors.  A
 *     non-positive Cholesky pivot (the reference's `assert`s on positive-definiteness,
 *     svae/lds/gaussian.py:48,64; LAPACK `info` is ignored by the Cython path,
 *     cython_gaussian_grads.pxd:54) is reported through the optional device word `status`
 *     (set to 1 + index of a failing sequence; left untouched otherwise).
 */
#ifndef SVAE_B200_H_
#define SVAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVAE_OK 0
#define SVAE_ERR_BAD_ARG (-1)       /* NULL pointer / non-positive size                           */
#define SVAE_ERR_UNSUPPORTED (-2)   /* n, K or num_samples outside the compiled range             */
#define SVAE_ERR_WORKSPACE (-3)     /* workspace too small                                         */
#define SVAE_ERR_CUDA (-4)          /* launch failed (cudaGetLastError != cudaSuccess)            */

#define SVAE_LDS_MAX_N 16           /* latent dimension range compiled into the library: 1..16    */
#define SVAE_HMM_MAX_K 32           /* discrete states of the register-resident kernels: 1..32 (lane = state) */
#define SVAE_HMM_WIDE_MAX_K 1024    /* svae_hmm_estep_wide / svae_hmm_viterbi: 1..1024 (thread = state)       */
#define SVAE_HMM_TC_MIN_K 17        /* svae_hmm_estep_wide: 17..32 states run on the fp64 tensor-core instruction */
#define SVAE_GMM_MAX_N 8            /* GMM latent dimension: 1..8                                  */

/* flags for svae_lds_estep_fwd */
#define SVAE_LDS_PER_T_PAIR_STATS 1 /* also write per-time second moments (inhomogeneous E_pair)   */
#define SVAE_LDS_SAVE_FOR_BWD 2     /* keep what svae_lds_estep_bwd needs in the workspace         */
#define SVAE_LDS_INIT_PER_SEQ 4     /* init_J (B,n,n), init_h (B,n), init_logZ (B): one per sequence */
#define SVAE_LDS_BWD_MESSAGES_OUT 32 /* svae_lds_estep_bwd_ports: adjoint of the smoother/sampler alone   */
#define SVAE_LDS_BWD_MESSAGES_IN 64  /* svae_lds_estep_bwd_ports: adjoint of the filter alone              */

/* Library version / build info: "svae_b200 <version> sm_100a". */
const char* svae_b200_version(void);

/* Kernels launched by this library since load (all entry points), for bench.py's gpu_launches. */
uint64_t svae_b200_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * LDS local E-step, forward: fused information-form Kalman filter + RTS smoother + backward sampler.
 * Replaces cython_natural_lds_inference_general (svae/lds/lds_inference.py:196-202), i.e. the
 * composition natural_filter_forward_general -> natural_smoother_general -> natural_sample_backward,
 * for B sequences at once.
 *
 *   init_J (n,n), init_h (n), init_logZ (1) or NULL          shared by the batch (per sequence with
 *       SVAE_LDS_INIT_PER_SEQ: the SLDS initial-state mixture, svae/models/slds_svae.py:92-103)
 *   J11, J12, J22 (n,n), pair_logZ (1) or NULL               homogeneous dynamics, shared; or, with
 *       pair_stride_t != 0: (T-1,n,n)/(T-1) time-inhomogeneous, element (t) at ptr + t*pair_stride_t
 *       (for pair_logZ the strides are divided by n*n); pair_stride_b != 0 adds a per-sequence stride
 *       (SLDS: svae/models/slds_svae.py:92-103).  Strides are in doubles.
 *   node_J, node_h (B,T,n), node_logZ (B,T) or NULL
 *   eps (B,T,S,n) standard-normal noise, eps[b,t] consumed at time t; NULL iff S == 0.
 *       (The reference draws randn(T,S,n) internally and consumes it in reverse time order,
 *        cython_lds_inference.pyx:333 -- draw k is used at time T-1-k.)
 * outputs (any may be NULL to skip):
 *   E_init_xx (B,n,n) E[x0 x0'], E_init_x (B,n) E[x0]
 *   E_pair_xx (B,n,n) sum_{t<T-1} E[x_t x_t'], E_pair_xxn (B,n,n) sum E[x_t x_{t+1}'],
 *   E_pair_xnxn (B,n,n) sum E[x_{t+1} x_{t+1}']           (cython_lds_inference.pyx:198-210)
 *   E_node_xx (B,T,n) diag E[x_t x_t'], E_node_x (B,T,n) E[x_t]
 *   samples (B,T,S,n)
 *   lognorm (B)
 *   Exx_t (B,T,n,n), Exxn_t (B,T-1,n,n): per-time moments, only with SVAE_LDS_PER_T_PAIR_STATS
 *   status: optional device int32
 *   workspace: device scratch of at least svae_lds_estep_workspace_bytes() bytes, 16-byte aligned.
 */
size_t svae_lds_estep_workspace_bytes(int B, int T, int n, int S, int flags);

int svae_lds_estep_fwd(
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
    int* status, void* workspace, size_t workspace_bytes, void* stream);

/* The same kernel behind the reference's three-primitive interface (svae/lds/lds_inference.py:18-39):
 *   msg_J_pred (B,T,n,n), msg_h_pred (B,T,n): OUTPUT the prediction messages of natural_filter_forward_general
 *       (cython_lds_inference.pyx:83-86; the filtered messages are J_pred + diag(node_J), h_pred + node_h);
 *   in_J_filt (B,T,n,n), in_h_filt (B,T,n): run the smoother / sampler sweeps from GIVEN filtered messages
 *       (natural_smoother_general, natural_sample_backward: pyx:149-195, 310-355) -- pass node_J = node_h = 0;
 *       `lognorm` is then meaningless;
 *   node_J_dense (B,T,n,n): dense node precisions, natural form (-1/2 J), ADDED to the diagonal node_J -- the
 *       reference's pure-Python option (svae/lds/lds_inference.py:65-80, svae/lds/gaussian.py:46-49) that its Cython
 *       path rejects (pyx:43).  With it E_node_stats[0] is the full E[x_t x_t'] = Exx_t (lds_inference.py:164-167):
 *       pass SVAE_LDS_PER_T_PAIR_STATS and Exx_t (Exxn_t may stay NULL).
 *   term_J (B,n,n): a dense precision (natural form) added at the LAST step only -- how a chunk of a longer chain
 *       receives the message from its future (parallel-in-time driver, svae_lds_compose_chunks below).
 * Any of the six may be NULL.
 */
int svae_lds_estep_fwd_ex(
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
    int* status, void* workspace, size_t workspace_bytes, void* stream);

/* svae_lds_estep_fwd_ex with a per-sequence mask: sequences with active[b] == 0 are skipped and every output row
 * of theirs keeps its previous content (active == NULL: all sequences run).  Used by the SLDS mean field, where each
 * sequence stops iterating once ITS bound has converged (svae/models/slds_svae.py:171-172) while the batch goes on. */
int svae_lds_estep_fwd_masked(
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
    int* status, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Parallel-in-time stage 1 (small batch, long T).  Cuts every sequence into
 * C = svae_lds_compose_num_chunks(T, chunk_len) chunks of chunk_len steps (the last may be shorter) and integrates the
 * node and pair potentials of each chunk -- the factors natural_filter_forward_general eliminates one by one
 * (svae/lds/lds_inference.py:86-106, svae/lds/gaussian.py:46-66) -- into ONE pair potential on the chunk's end states:
 *   chunk_J11, chunk_J12, chunk_J22 (B,C,n,n), chunk_logZ (B,C): natural form, the layout svae_lds_estep_fwd takes as
 *       per-sequence time-inhomogeneous pair parameters (pair_stride_t = n*n, pair_stride_b = C*n*n);
 *   chunk_e1, chunk_e2 (B,C,n): linear terms on the chunk's first / last state (node_h of the boundary chain).
 * Pair parameters addressed like svae_lds_estep_fwd (pair_stride_t = pair_stride_b = 0: homogeneous (n,n) blocks and
 * one pair_logZ; else per-time / per-sequence blocks with pair_logZ indexed by stride / n^2); diagonal node potentials
 * (B,T,n); node_logZ (B,T) / pair_logZ may be NULL.
 * *status (may be NULL): 1 + index of the first sequence with a non-positive pivot.
 */
int svae_lds_compose_num_chunks(int T, int chunk_len);
int svae_lds_compose_chunks(int B, int T, int n, int chunk_len, const double* J11, const double* J12,
                            const double* J22, const double* pair_logZ, long long pair_stride_t,
                            long long pair_stride_b, const double* node_J,
                            const double* node_h, const double* node_logZ, double* chunk_J11,
                            double* chunk_J12, double* chunk_J22, double* chunk_logZ, double* chunk_e1,
                            double* chunk_e2, int* status, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LDS local E-step, backward: the VJP of svae_lds_estep_fwd with respect to the node potentials
 * (node_J, node_h), given cotangents of every output.  Replaces the composition
 *   natural_sample_backward_grad + natural_smoother_general_grad -> natural_filter_grad
 * (svae/lds/cython_lds_inference.pyx:357-409, 236-306, 92-145) as wired at svae/lds/lds_inference.py:26-39;
 * like the reference, gradients flow to the node parameters only (lds_inference.py:27-29).
 *
 * The forward call must have been made with SVAE_LDS_SAVE_FOR_BWD; `workspace` is that call's workspace
 * (it is consumed: single use, as the reference's `intermediates` are -- pyx:472,481), `E_node_x`,
 * `samples` and `eps` are that call's outputs / noise.
 *   cotangents (device pointers, any may be NULL = zero), shaped like the corresponding outputs:
 *     g_samples (B,T,S,n); g_E_init_xx (B,n,n), g_E_init_x (B,n);
 *     g_E_pair_xx, g_E_pair_xxn, g_E_pair_xnxn (B,n,n)          [homogeneous sums]
 *     g_Exx_t (B,T,n,n), g_Exxn_t (B,T-1,n,n)                   [per-time moments]
 *     g_E_node_xx, g_E_node_x (B,T,n); g_lognorm (B)
 *   outputs: g_node_J, g_node_h (B,T,n).  (d lognorm / d node_logZ is g_lognorm itself.)
 */
int svae_lds_estep_bwd(
    int B, int T, int n, int S, int flags,
    const double* E_node_x, const double* samples, const double* eps,
    const double* g_samples,
    const double* g_E_init_xx, const double* g_E_init_x,
    const double* g_E_pair_xx, const double* g_E_pair_xxn, const double* g_E_pair_xnxn,
    const double* g_Exx_t, const double* g_Exxn_t,
    const double* g_E_node_xx, const double* g_E_node_x, const double* g_lognorm,
    double* g_node_J, double* g_node_h,
    void* workspace, size_t workspace_bytes, void* stream);

/* As above plus g_node_J_dense (B,T,n,n): the un-symmetrised cotangent Pbar_t of the filtered precision at every
 * step; the gradient w.r.t. a dense natural-form node precision is -(Pbar_t + Pbar_t') (g_node_J may then be NULL).
 * In this entry point the homogeneous-sum cotangents (g_E_pair_*) and the per-time ones (g_Exx_t, g_Exxn_t) ADD. */
int svae_lds_estep_bwd_ex(
    int B, int T, int n, int S, int flags,
    const double* E_node_x, const double* samples, const double* eps,
    const double* g_samples,
    const double* g_E_init_xx, const double* g_E_init_x,
    const double* g_E_pair_xx, const double* g_E_pair_xxn, const double* g_E_pair_xnxn,
    const double* g_Exx_t, const double* g_Exxn_t,
    const double* g_E_node_xx, const double* g_E_node_x, const double* g_lognorm,
    double* g_node_J, double* g_node_h, double* g_node_J_dense,
    void* workspace, size_t workspace_bytes, void* stream);

/* The reference differentiates its three primitives ONE BY ONE (primitive_with_aux, svae/lds/lds_inference.py:26-39):
 * natural_smoother_general_grad and natural_sample_backward_grad return cotangents w.r.t. the forward MESSAGES
 * (cython_lds_inference.pyx:236-306, 357-409), natural_filter_grad takes cotangents of the messages and of lognorm and
 * returns those of the node potentials (pyx:92-145).  svae_lds_estep_bwd_ports is svae_lds_estep_bwd_ex with those ports:
 *   flags | SVAE_LDS_BWD_MESSAGES_OUT : run the adjoint of the smoother + sampler only; g_msg_M (B,T,n,n), g_msg_h (B,T,n)
 *       receive the cotangents of M_t = (filtered precision + J11, information form: filtered natural J = -1/2 (M_t - J11))
 *       and of the filtered h_t; the covariance-form smoother never reads the prediction messages, whose cotangent is zero
 *       (the reference's information-form RTS step splits the same total between the two; the totals pushed through
 *       natural_filter_grad agree, tests/test_lds_primitives_gpu.py).  g_node_* may be NULL.
 *   flags | SVAE_LDS_BWD_MESSAGES_IN  : run the adjoint of the filter only; in_g_M_pred / in_g_M_filt (B,T,n,n) and
 *       in_g_h_pred / in_g_h_filt (B,T,n) are cotangents of the prediction / filtered messages (information form: of
 *       J_info = -2 J_natural, i.e. -1/2 x the natural-form cotangent, symmetrised by the caller; any may be NULL),
 *       g_lognorm that of the log-normaliser; outputs g_node_J / g_node_h (/ g_node_J_dense).
 * The workspace is the one a forward launch with SVAE_LDS_SAVE_FOR_BWD filled (with or without message ports). */
int svae_lds_estep_bwd_ports(
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
    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Minibatch reduction of the LDS global expected statistics: out[4n^2+n+1] =
 *   sum_b [ E_init_xx | E_init_x | E_pair_xx | E_pair_xxn | E_pair_xnxn | lognorm ]   (deterministic order).
 * This is flat(stats) of svae/svae.py:33-34 for a minibatch of B sequences, and the payload of the single
 * all-reduce per SVI step when the minibatch is sharded over GPUs.
 */
size_t svae_lds_reduce_stats_workspace_bytes(int n);
int svae_lds_reduce_stats(int B, int n, const double* E_init_xx, const double* E_init_x,
                          const double* E_pair_xx, const double* E_pair_xxn, const double* E_pair_xnxn,
                          const double* lognorm, double* out, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * HMM local E-step for B sequences: log-normaliser and expected statistics by scaled forward-backward.
 * Replaces hmm_logZ / hmm_logZ_grad (svae/hmm/cython_hmm_inference.pyx:93-166; with cotangent 1 the
 * gradient IS the E-step, svae/hmm/hmm_inference.py:65) and hmm_estep (hmm_inference.py:21-41, pyhsmm).
 *   log_pi (K), log_A (K,K) row = from-state  (unnormalised log values allowed), shared by the batch
 *   node (B,T,K) log node potentials
 * outputs: logZ (B); Ez0 (B,K) = E[z_0]; Etrans (B,K,K) expected transition counts; Ez (B,T,K) = E[z_t].
 * With SVAE_HMM_LOGZ_ONLY only logZ is written (Ez0, Etrans, Ez may be NULL); otherwise Ez is required
 * (it doubles as the scratch for the scaled forward messages).  The VJP of logZ w.r.t. (log_pi, log_A,
 * node) is cotangent x (Ez0, Etrans, Ez) -- exactly hmm_logZ_grad (pyx:126-166).
 */
#define SVAE_HMM_LOGZ_ONLY 1
int svae_hmm_estep(int B, int T, int K, int flags, const double* log_pi, const double* log_A,
                   const double* node, double* logZ, double* Ez0, double* Etrans, double* Ez,
                   void* stream);

/* svae_hmm_estep with a per-sequence mask (see svae_lds_estep_fwd_masked). */
int svae_hmm_estep_masked(int B, int T, int K, int flags, const double* log_pi, const double* log_A,
                          const double* node, const int* active, double* logZ, double* Ez0, double* Etrans,
                          double* Ez, void* stream);

/* The same E-step for ANY number of states up to SVAE_HMM_WIDE_MAX_K (the reference's hmm_logZ / hmm_logZ_grad take any
 * K, svae/hmm/cython_hmm_inference.pyx:93-166) -- the entry point a caller with a scratch buffer should use:
 *   K <  SVAE_HMM_TC_MIN_K        the register-resident kernel of svae_hmm_estep_masked (workspace ignored);
 *   SVAE_HMM_TC_MIN_K..32 states  with a workspace: 8 sequences per warp, the K x K products and the transition counts on
 *                                 the fp64 tensor-core instruction (csrc/hmm_tc.cu; the workspace holds the scaled forward
 *                                 messages and w_{t+1}/s_t, 2 B T K doubles; not needed with SVAE_HMM_LOGZ_ONLY);
 *                                 workspace NULL otherwise: as above;
 *   33..SVAE_HMM_WIDE_MAX_K       one CTA per sequence, thread = state (csrc/hmm_wide.cu); the workspace (required) holds
 *                                 exp(log_A - max), its transpose and the shift, written by a one-CTA preparation launch.
 * Same arguments, outputs and mask otherwise; everything is enqueued on `stream`. */
size_t svae_hmm_wide_workspace_bytes(int B, int T, int K);
int svae_hmm_estep_wide(int B, int T, int K, int flags, const double* log_pi, const double* log_A,
                        const double* node, const int* active, double* logZ, double* Ez0, double* Etrans,
                        double* Ez, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SLDS local mean field: the two contractions that couple the HMM factor and the LDS factor, batched over
 * (sequence, time).  K discrete modes with per-mode parameters stacked on a leading K axis:
 *   init_J (K,n,n), init_h (K,n), init_c (K) [= the two constants of the packed NIW statistics added], and
 *   J11, J12, J22 (K,n,n), pair_logZ (K): the expected MNIW statistics, i.e. LDS pair potentials per mode.
 * svae_slds_mix_params   (svae/models/slds_svae.py:92-103): out_init_* (B,..) = sum_k Ez[b,0,k] init_k,
 *                        out_J11/J12/J22 (B,T-1,n,n), out_pair_logZ (B,T-1): at t the mix with Ez[b,t+1,k].
 * svae_slds_node_potentials (slds_svae.py:131-147): node (B,T,K), node[b,0,k] = <E[x0 x0'], init_J_k> + <E[x0], init_h_k>
 *                        + init_c_k;  node[b,t,k] = <E[x_{t-1}x_{t-1}'], J11_k> + <E[x_{t-1}x_t'], J12_k> + <E[x_t x_t'], J22_k>
 *                        + pair_logZ_k, from the per-time second moments Exx_t (B,T,n,n), Exxn_t (B,T-1,n,n) that
 *                        svae_lds_estep_fwd writes with SVAE_LDS_PER_T_PAIR_STATS, and Ex0 (B,n) = E[x_0].
 * active: optional (B,) mask as in svae_lds_estep_fwd_masked.  K <= 32; returns SVAE_ERR_UNSUPPORTED when the K parameter
 * blocks do not fit in shared memory (4 K n^2 doubles). */
int svae_slds_mix_params(int B, int T, int n, int K, const double* Ez, const double* init_J, const double* init_h,
                         const double* init_c, const double* J11, const double* J12, const double* J22,
                         const double* pair_logZ, const int* active, double* out_init_J, double* out_init_h,
                         double* out_init_c, double* out_J11, double* out_J12, double* out_J22, double* out_pair_logZ,
                         void* stream);
int svae_slds_node_potentials(int B, int T, int n, int K, const double* Exx_t, const double* Exxn_t, const double* Ex0,
                              const double* init_J, const double* init_h, const double* init_c, const double* J11,
                              const double* J12, const double* J22, const double* pair_logZ, const int* active,
                              double* node, void* stream);

/* Parallel-in-time HMM E-step, stages 1 and 2 (the sequential recursion is
 * svae/hmm/hmm_inference.py:44-51, its reverse sweep svae/hmm/cython_hmm_inference.pyx:126-166).
 *   svae_hmm_compose_chunks: C = svae_lds_compose_num_chunks(T, chunk_len) chunks per sequence; chunk_P (B,C,K,K) =
 *       prod_t diag(exp(node_t)) A over the chunk's steps, scaled to max 1, chunk_logscale (B,C) the log of the scale.
 *   svae_hmm_chain_messages: forward / backward messages at the C+1 chunk boundaries, log_alpha, log_beta (B,C+1,K)
 *       (log_alpha excludes, log_beta includes the node potential at the boundary; node_last (B,K) is the node
 *       potential of the last time step), and the total log-normaliser logZ (B).
 * Stage 3 is svae_hmm_estep on the B*C chunks with log_alpha added to each chunk's first node potential, log_beta
 * replacing its last one and log_pi = 0 (svae_b200/hmm/hmm_pit.py). */
int svae_hmm_compose_chunks(int B, int T, int K, int chunk_len, const double* log_A, const double* node,
                            double* chunk_P, double* chunk_logscale, void* stream);
int svae_hmm_chain_messages(int B, int C, int K, const double* log_pi, const double* chunk_P,
                            const double* chunk_logscale, const double* node_last, double* log_alpha,
                            double* log_beta, double* logZ, void* stream);

/* Max-plus twins for the most likely state path (svae/hmm/hmm_inference.py:54-63) of few long sequences:
 *   svae_hmm_compose_chunks_maxplus: chunk_V (B,C,K,K), V(i,j) = best summed log-potential of a path from z_s = i to
 *       z_{s+L} = j through the chunk (node potentials s .. s+L-1 included);
 *   svae_hmm_chain_viterbi: max-product recursion over the chunk matrices and back-tracking: boundary_states (B,C+1)
 *       int32 = the optimal state at every chunk boundary (first maximal index on ties); delta_scratch (B,C+1,K).
 * Stage 3 is svae_hmm_viterbi on the B*C chunks with both end states pinned by -1e300 penalties.  The phase code of both
 * is pinned on CPU (tests/test_hmm_pit_cpu.py) and checked on the GPU against the sequential kernel (tests/test_hmm_pit_gpu.py). */
int svae_hmm_compose_chunks_maxplus(int B, int T, int K, int chunk_len, const double* log_A, const double* node,
                                    double* chunk_V, void* stream);
int svae_hmm_chain_viterbi(int B, int C, int K, const double* log_pi, const double* chunk_V,
                           const double* node_last, double* delta_scratch, int* boundary_states, void* stream);

/* Viterbi path (svae/hmm/hmm_inference.py:54-63, pyhsmm `viterbi`): path (B,T) int32, first maximal
 * index on ties.  workspace: svae_hmm_viterbi_workspace_bytes() bytes of device scratch (back-pointers: one byte per
 * (b,t,k) for K <= SVAE_HMM_MAX_K, two for K up to SVAE_HMM_WIDE_MAX_K, which runs one CTA per sequence). */
size_t svae_hmm_viterbi_workspace_bytes(int B, int T, int K);
int svae_hmm_viterbi(int B, int T, int K, const double* log_pi, const double* log_A, const double* node,
                     int32_t* path, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * GMM-SVAE local mean-field fixed point (svae/models/gmm.py:90-110 meanfield_fixed_point with
 * gaussian_meanfield :112-117 and label_meanfield :119-124) for B independent minibatches of T points,
 * the whole coordinate ascent inside one launch (one CTA per minibatch, no host round trip per iteration).
 *   node_J, node_h (B,T,N): diagonal Gaussian node potentials (-1/2 precision, h)      [gaussian_info heads]
 *   label_global (K) = E[log pi]  (dirichlet.expectedstats); gaussian_globals (K,N+2,N+2) = packed
 *   niw.expectedstats (svae/distributions/gaussian.py:39-53 layout)
 *   label_init (B,T,K): initial responsibilities (the reference draws them with npr.rand, gmm.py:126-128)
 *   tol, max_iter: 1e-3 and 100 in the reference
 * outputs: label_stats (B,T,K) responsibilities at the fixed point; iters (B) iterations used; kl (B) the
 * local KL of the last sweep.  The differentiable evaluation at the optimum stays with the caller.
 */
int svae_gmm_meanfield(int B, int T, int K, int N, const double* node_J, const double* node_h,
                       const double* label_global, const double* gaussian_globals,
                       const double* label_init, double tol, int max_iter, double* label_stats,
                       int* iters, double* kl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The recognition network's output head fused with node-potential production (svae/nnet.py:43-47 gaussian_info and the
 * 'info' branch of the Gaussian resnet, nnet.py:90-104):
 *   act (rows, 2n) = [J_in | h_in] last-layer activations;  res (rows, n) = inputs W + b1 or NULL;  b2 (n) or NULL
 *   node_J = -1/2 (log1pexp(J_in) + log1pexp(b2)),  node_h = h_in + res        (rows = B*T of the E-step's (B,T,n))
 * and its VJP w.r.t. act: g_act (rows, 2n) = [-1/2 sigmoid(J_in) g_node_J | g_node_h]   (d/d res = g_node_h,
 * d/d b2 = -1/2 sigmoid(b2) sum_rows g_node_J are formed by the caller). */
int svae_gaussian_info_head_fwd(long long rows, int n, const double* act, const double* res, const double* b2,
                                double* node_J, double* node_h, void* stream);
int svae_gaussian_info_head_bwd(long long rows, int n, const double* act, const double* g_node_J, const double* g_node_h,
                                double* g_act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVAE_B200_H_ */
