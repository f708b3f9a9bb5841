// Dev microbenchmark: what does a shared-memory load cost on B200 as a function of the lane -> address
// pattern (broadcast groups) and the access width?  Cycles per warp-instruction at the SM level, many warps
// resident, loads independent.  Also: DMMA (fp64 tensor) vs DFMA throughput, and whether they share a pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bcast lds_bcast.cu && ./lds_bcast
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int W>   // bytes per lane: 4, 8, 16
__global__ void lds_tput(const int* lane_off, int iters, long long* cyc, unsigned* sink, int active_mask_lanes) {
  extern __shared__ __align__(16) unsigned char sm[];
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + lane_off[lane] + (threadIdx.x >> 5) * 1024 % 8192;
  unsigned acc = 0;
  const bool act = lane < active_mask_lanes;
  long long t0 = clock64();
  if (act) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        unsigned a, b, c, d;
        if (W == 16) {
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + j * 256 + ((it & 1) << 12)) : "memory");
          acc ^= a;
        } else if (W == 8) {
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(base + j * 256 + ((it & 1) << 12)) : "memory");
          acc ^= a;
        } else {
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(base + j * 256 + ((it & 1) << 12)) : "memory");
          acc ^= a;
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int W>
__global__ void sts_tput(const int* lane_off, int iters, long long* cyc, unsigned* sink) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int lane = threadIdx.x & 31;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + lane_off[lane] + (threadIdx.x >> 5) * 1024 % 8192;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (W == 8) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(base + j * 256 + ((it & 1) << 12)), "r"(it), "r"(j) : "memory");
      else asm volatile("st.shared.v4.u32 [%0], {%1,%2,%1,%2};" ::"r"(base + j * 256 + ((it & 1) << 12)), "r"(it), "r"(j) : "memory");
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}

// DFMA throughput (ILP 8) optionally mixed with DMMA m8n8k4
template <int NF, int NM>
__global__ void fp64_mix(int iters, long long* cyc, double* sink) {
  double a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3 + j;
  double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  double fa = 1.0 + threadIdx.x * 1e-9, fb = 1.0 - threadIdx.x * 1e-9;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < NF; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fma(a[j], b, c);
#pragma unroll
    for (int r = 0; r < NM; ++r) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[r & 3]), "+d"(c1[r & 3]) : "d"(fa), "d"(fb));
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  double s = 0;
  for (int j = 0; j < 8; ++j) s += a[j];
  for (int j = 0; j < 4; ++j) s += c0[j] + c1[j];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA fed by broadcast LDS.128: 2 FMA per loaded double2 (the smoother's inner loop shape), R accumulator sets
template <int R>
__global__ void lds_fma(const int* lane_off, int iters, long long* cyc, double* sink) {
  extern __shared__ __align__(16) unsigned char sm[];
  for (int i = threadIdx.x; i < 16384 / 8; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = 1e-3 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + lane_off[lane] + (threadIdx.x >> 5) * 1024 % 8192;
  double acc[R][10];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[r][j] = 0;
  double z[R];
#pragma unroll
  for (int r = 0; r < R; ++r) z[r] = 1.0 + lane * 1e-3 + r;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        double x, y;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(base + (k * 5 + q) * 16 + ((it & 1) << 12)) : "memory");
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[r][2 * q] = fma(x, z[r], acc[r][2 * q]);
          acc[r][2 * q + 1] = fma(y, z[r], acc[r][2 * q + 1]);
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  double s = 0;
  for (int r = 0; r < R; ++r)
    for (int j = 0; j < 10; ++j) s += acc[r][j];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static long long maxcyc(long long* d, int n) {
  static long long h[4096];
  cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost);
  long long m = 0;
  for (int i = 0; i < n; ++i) m = h[i] > m ? h[i] : m;
  return m;
}

int main() {
  cudaError_t e;
  int nsm = 148;
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  nsm = pr.multiProcessorCount;
  printf("device %s, %d SMs\n", pr.name, nsm);
  long long* cyc;
  unsigned* sink;
  int* off;
  cudaMalloc(&cyc, 4096 * sizeof(long long));
  cudaMalloc(&sink, 4096 * 1024 * 8);
  cudaMalloc(&off, 32 * sizeof(int));
  struct Pat { const char* name; int off[32]; } pats[16];
  int np = 0;
  auto mk = [&](const char* name, auto f) { pats[np].name = name; for (int l = 0; l < 32; ++l) pats[np].off[l] = f(l); ++np; };
  mk("all lanes same address", [](int l) { return 0; });
  mk("2 halves, A / A+4656 (today's tiles)", [](int l) { return (l / 16) * 4656; });
  mk("2 halves, A / A+16 (same 128B line)", [](int l) { return (l / 16) * 16; });
  mk("2 halves, A / A+128", [](int l) { return (l / 16) * 128; });
  mk("4 quarters, stride 4656", [](int l) { return (l / 8) * 4656 % 12288; });
  mk("4 quarters, stride 16 (one line)", [](int l) { return (l / 8) * 16; });
  mk("3 groups of 10 (+2), stride 4656", [](int l) { int g = l / 10; if (g > 2) g = 2; return g * 4656 % 12288; });
  mk("3 groups of 10 (+2), stride 16", [](int l) { int g = l / 10; if (g > 2) g = 2; return g * 16; });
  mk("8 groups of 4, stride 16", [](int l) { return (l / 4) * 16; });
  mk("8 groups of 4, stride 528", [](int l) { return (l / 4) * 528; });
  mk("interleaved: lane%2 -> A / A+16", [](int l) { return (l % 2) * 16; });
  mk("all distinct, contiguous 16B", [](int l) { return l * 16; });
  mk("all distinct, contiguous 8B", [](int l) { return l * 8; });
  mk("column store: stride 80B (NP=10)", [](int l) { return (l % 16) * 80 + (l / 16) * 4656; });
  mk("column store: stride 88B (NP=11)", [](int l) { return (l % 16) * 88 + (l / 16) * 4656; });
  mk("column store: stride 96B (NP=12)", [](int l) { return (l % 16) * 96 + (l / 16) * 4656; });
  const int iters = 2000, warps = 8, blocks_per_sm = 2;
  const int grid = nsm * blocks_per_sm;
  const double per = (double)iters * 16 * warps * blocks_per_sm;   // warp-instructions per SM
  printf("\n== LDS throughput: SM cycles per warp-instruction (%d warps/SM resident) ==\n", warps * blocks_per_sm);
  printf("%-42s %8s %8s %8s   %s\n", "pattern", "LDS.32", "LDS.64", "LDS.128", "STS.64 / STS.128");
  for (int p = 0; p < np; ++p) {
    cudaMemcpy(off, pats[p].off, sizeof(int) * 32, cudaMemcpyHostToDevice);
    double r[5];
    int al = 16;
    for (int l = 0; l < 32; ++l) { if (pats[p].off[l] % 16) al = 8; }
    for (int l = 0; l < 32; ++l) { if (pats[p].off[l] % 8) al = 4; }
    r[0] = r[1] = r[2] = r[3] = r[4] = -1;
    lds_tput<4><<<grid, warps * 32, 40960>>>(off, iters, cyc, sink, 32); cudaDeviceSynchronize(); r[0] = maxcyc(cyc, grid) / per;
    if (al >= 8) { lds_tput<8><<<grid, warps * 32, 40960>>>(off, iters, cyc, sink, 32); cudaDeviceSynchronize(); r[1] = maxcyc(cyc, grid) / per; }
    if (al >= 16) { lds_tput<16><<<grid, warps * 32, 40960>>>(off, iters, cyc, sink, 32); cudaDeviceSynchronize(); r[2] = maxcyc(cyc, grid) / per; }
    if (al >= 8) { sts_tput<8><<<grid, warps * 32, 40960>>>(off, iters, cyc, sink); cudaDeviceSynchronize(); r[3] = maxcyc(cyc, grid) / per; }
    if (al >= 16) { sts_tput<16><<<grid, warps * 32, 40960>>>(off, iters, cyc, sink); cudaDeviceSynchronize(); r[4] = maxcyc(cyc, grid) / per; }
    printf("%-42s %8.2f %8.2f %8.2f   %.2f / %.2f\n", pats[p].name, r[0], r[1], r[2], r[3], r[4]);
  }
  // predicated-off lanes: 12 of 16 active per half
  {
    int o[32];
    for (int l = 0; l < 32; ++l) o[l] = (l / 16) * 4656;
    cudaMemcpy(off, o, sizeof(o), cudaMemcpyHostToDevice);
    for (int act : {32, 24, 16, 8}) {
      lds_tput<16><<<grid, warps * 32, 40960>>>(off, iters, cyc, sink, act); cudaDeviceSynchronize();
      printf("LDS.128 2 halves, first %d lanes active: %.2f\n", act, maxcyc(cyc, grid) / per);
    }
  }
  printf("\n== fp64: SM cycles per warp-instruction (16 warps/SM) ==\n");
  {
    double* ds = reinterpret_cast<double*>(sink);
    const int it2 = 2000;
    fp64_mix<4, 0><<<grid, 256>>>(it2, cyc, ds); cudaDeviceSynchronize();
    printf("DFMA only              : %.3f cycles per warp-DFMA\n", maxcyc(cyc, grid) / ((double)it2 * 32 * 16));
    fp64_mix<0, 8><<<grid, 256>>>(it2, cyc, ds); cudaDeviceSynchronize();
    printf("DMMA m8n8k4 only       : %.3f cycles per warp-DMMA (256 FMA each)\n", maxcyc(cyc, grid) / ((double)it2 * 8 * 16));
    fp64_mix<4, 8><<<grid, 256>>>(it2, cyc, ds); cudaDeviceSynchronize();
    printf("32 DFMA + 8 DMMA mixed : %.1f cycles per iteration per SM-warp (sum if shared pipe: see above)\n",
           maxcyc(cyc, grid) / ((double)it2 * 16));
  }
  printf("\n== broadcast LDS.128 feeding DFMA (smoother inner-loop shape): SM cycles per LDS.128 ==\n");
  {
    double* ds = reinterpret_cast<double*>(sink);
    for (int p : {0, 1, 6}) {
      cudaMemcpy(off, pats[p].off, sizeof(int) * 32, cudaMemcpyHostToDevice);
      const int it2 = 2000;
      lds_fma<1><<<grid, 256, 40960>>>(off, it2, cyc, ds); cudaDeviceSynchronize();
      double a = maxcyc(cyc, grid) / ((double)it2 * 20 * 16);
      lds_fma<2><<<grid, 256, 40960>>>(off, it2, cyc, ds); cudaDeviceSynchronize();
      double b = maxcyc(cyc, grid) / ((double)it2 * 20 * 16);
      lds_fma<3><<<grid, 256, 40960>>>(off, it2, cyc, ds); cudaDeviceSynchronize();
      double c = maxcyc(cyc, grid) / ((double)it2 * 20 * 16);
      printf("%-42s R=1 (2 FMA/ld): %.2f   R=2 (4 FMA/ld): %.2f   R=3 (6 FMA/ld): %.2f\n", pats[p].name, a, b, c);
    }
  }
  e = cudaDeviceSynchronize();
  printf("final status: %s\n", cudaGetErrorString(e));
  return 0;
}
