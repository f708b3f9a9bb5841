// One translation unit per latent dimension: compiled with -DSVAE_N=<n> for n = 1..SVAE_LDS_MAX_N so
// the fully-unrolled kernels build in parallel.  Exposes svae::lds_fwd_launch_n<n>().
#include "lds_estep.cuh"
#include "lds_estep_packed.cuh"
#include "launch.h"

#ifndef SVAE_N
#error "compile with -DSVAE_N=<latent dim>"
#endif

#define SVAE_CAT_(a, b) a##b
#define SVAE_CAT(a, b) SVAE_CAT_(a, b)

#ifdef SVAE_STUB
// experiment builds (SVAE_BUILD_NS=... python -m svae_b200.build): latent dimensions left out of a variant library
namespace svae {
cudaError_t SVAE_CAT(lds_fwd_launch_n, SVAE_N)(const LdsFwdParams&, int, int, bool, cudaStream_t) { return cudaErrorNotSupported; }
}
#else

#include <mutex>
#include <cstdlib>

namespace svae {

// The homogeneous J12 is handed to the kernel through this translation unit's __constant__ slots.
// Slots rotate; a slot is re-used from a different stream only after the event recorded behind its last
// kernel, so a copy never overwrites a bank another stream's launch may still read.
static std::mutex g_const_mu;
constexpr int kMaxDevices = 32;     // per-device state: __constant__ symbols and events belong to one device
struct ConstSlots {
  cudaEvent_t ev[SVAE_LDS_CONST_SLOTS] = {};
  cudaStream_t stream[SVAE_LDS_CONST_SLOTS] = {};
  bool used[SVAE_LDS_CONST_SLOTS] = {};
  int next = 0;
};
static ConstSlots g_const[kMaxDevices];

template <int N>
static cudaError_t stage_const_J12(const double* J12, cudaStream_t stream, int* slot_out, ConstSlots** st_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  ConstSlots& cs = g_const[dev];
#ifdef SVAE_PACKED_SLOT0
  const int slot = 0;
#else
  const int slot = cs.next;
  cs.next = (cs.next + 1) % SVAE_LDS_CONST_SLOTS;
#endif
  if (!cs.ev[slot]) {
    e = cudaEventCreateWithFlags(&cs.ev[slot], cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  if (cs.used[slot] && cs.stream[slot] != stream) {
    e = cudaStreamWaitEvent(stream, cs.ev[slot], 0);
    if (e != cudaSuccess) return e;
  }
  *slot_out = slot;
  *st_out = &cs;
  return cudaMemcpyToSymbolAsync(c_J12, J12, sizeof(double) * N * N, sizeof(double) * SVAE_LDS_CONST_NN * slot,
                                 cudaMemcpyDeviceToDevice, stream);
}

template <int N, int CPL, int LPS, bool INHOMOG>
static cudaError_t launch_one(const LdsFwdParams& p, cudaStream_t stream) {
  constexpr int SPW = 32 / LPS;
  const int warps_per_cta = 1;
  const long long tiles = (p.B + SPW - 1) / SPW;
  const long long grid = (tiles + warps_per_cta - 1) / warps_per_cta;
  const size_t smem = sizeof(double) *
      (3 * N * lds_np(N) + (size_t)warps_per_cta * SPW * lds_tile_doubles(N, p.S, INHOMOG));
  auto kern = lds_estep_fwd_kernel<N, CPL, LPS, INHOMOG>;
  {
    // many small CTAs per SM: optional shared-memory carve-out hint (percent), SVAE_LDS_CARVEOUT
    static bool carved = false;
    if (!carved) {
      const char* e = getenv("SVAE_LDS_CARVEOUT");
      if (e) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
      carved = true;
    }
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (!INHOMOG && p.T > 1) {
    std::lock_guard<std::mutex> lock(g_const_mu);
    LdsFwdParams q = p;
    int slot = 0;
    ConstSlots* cs = nullptr;
    cudaError_t e = stage_const_J12<N>(p.J12, stream, &slot, &cs);
    if (e != cudaSuccess) return e;
    q.const_slot = slot;
    kern<<<(unsigned)grid, warps_per_cta * 32, smem, stream>>>(q);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    cs->used[slot] = true;
    cs->stream[slot] = stream;
    return cudaEventRecord(cs->ev[slot], stream);
  }
  kern<<<(unsigned)grid, warps_per_cta * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

// the packed mapping (lds_estep_packed.cuh): N lanes per sequence, 32/N sequences per warp
template <int N, int S, bool SAVE, bool INHOMOG>
static cudaError_t launch_packed(const LdsFwdParams& p, cudaStream_t stream) {
  constexpr int SPW = lds_packed_spw(N);
  const long long grid = ((long long)p.B + SPW - 1) / SPW;
  const size_t smem = sizeof(double) * (3 * N * lds_np(N) + 2 + (size_t)SPW * lds_packed_tile_doubles(N, S, INHOMOG, lds_packed_nog(SAVE, INHOMOG)));
  auto kern = lds_estep_packed_kernel<N, S, SAVE, INHOMOG>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (p.T > 1 && !INHOMOG) {
    std::lock_guard<std::mutex> lock(g_const_mu);
    LdsFwdParams q = p;
    int slot = 0;
    ConstSlots* cs = nullptr;
    cudaError_t e = stage_const_J12<N>(p.J12, stream, &slot, &cs);
    if (e != cudaSuccess) return e;
    q.const_slot = slot;
    kern<<<(unsigned)grid, 32, smem, stream>>>(q);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    cs->used[slot] = true;
    cs->stream[slot] = stream;
    return cudaEventRecord(cs->ev[slot], stream);
  }
  kern<<<(unsigned)grid, 32, smem, stream>>>(p);
  return cudaGetLastError();
}

// the packed kernel covers the hot configuration only; everything else runs the general kernel
static bool packed_eligible(const LdsFwdParams& p, bool inhomog) {
  static const bool enabled = [] { const char* e = getenv("SVAE_LDS_PACKED"); return !(e && atoi(e) == 0); }();
  if (!enabled || p.S > 2 || p.msg_J_pred || p.msg_h_pred || p.in_J_filt || p.in_h_filt || p.node_J_dense || p.term_J ||
      (p.S > 0 && !p.samples))
    return false;
  // statistic buffers are optional, in pairs (natural_lds_sample passes no E_init_* / E_pair_*)
  if ((p.E_node_xx != nullptr) != (p.E_node_x != nullptr)) return false;
  if (inhomog) return true;       // per-sequence / per-time parameters, optional per-time outputs
  return !(p.flags & SVAE_LDS_FLAG_INIT_PER_SEQ) && !p.Exx_t && !p.Exxn_t;
}

template <int N, int CPL, int LPS>
static cudaError_t launch_if_fits(const LdsFwdParams& p, bool inhomog, cudaStream_t stream) {
  if constexpr (CPL * LPS >= N + 1) {
    return inhomog ? launch_one<N, CPL, LPS, true>(p, stream) : launch_one<N, CPL, LPS, false>(p, stream);
  } else {
    return cudaErrorInvalidValue;
  }
}

// cpl: columns per lane (1 or 2); lps: lanes per sequence tile, with cpl * lps >= n + 1 + S
cudaError_t SVAE_CAT(lds_fwd_launch_n, SVAE_N)(const LdsFwdParams& p, int cpl, int lps, bool inhomog, cudaStream_t stream) {
  constexpr int N = SVAE_N;
  if (packed_eligible(p, inhomog)) {
    const bool save = (p.flags & SVAE_LDS_FLAG_SAVE) != 0;
#define SVAE_PACKED_CASE(SS) \
    return inhomog ? (save ? launch_packed<N, SS, true, true>(p, stream) : launch_packed<N, SS, false, true>(p, stream)) \
                   : (save ? launch_packed<N, SS, true, false>(p, stream) : launch_packed<N, SS, false, false>(p, stream))
    switch (p.S) {
      case 0: SVAE_PACKED_CASE(0);
      case 1: SVAE_PACKED_CASE(1);
      default: SVAE_PACKED_CASE(2);
    }
#undef SVAE_PACKED_CASE
  }
#ifdef SVAE_LDS_ENABLE_CPL2   // two column slots per lane: measured slower than CPL = 1 on B200
  if (cpl == 2) {
    switch (lps) {
      case 4: return launch_if_fits<N, 2, 4>(p, inhomog, stream);
      case 8: return launch_if_fits<N, 2, 8>(p, inhomog, stream);
      case 16: return launch_if_fits<N, 2, 16>(p, inhomog, stream);
      default: return launch_if_fits<N, 2, 32>(p, inhomog, stream);
    }
  }
#else
  if (cpl == 2) return cudaErrorInvalidValue;
#endif
  if (lps == 16) return launch_if_fits<N, 1, 16>(p, inhomog, stream);
  return launch_if_fits<N, 1, 32>(p, inhomog, stream);
}

}  // namespace svae
#endif  // SVAE_STUB
