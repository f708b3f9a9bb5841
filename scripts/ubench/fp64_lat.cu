// Dev microbenchmark: fp64 pipe latency / throughput, smem round trip, on the actual B200.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dep_chain(double* out, int iters, long long* cyc) {
  double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a = fma(a, b, c);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
template <int ILP>
__global__ void tput(double* out, int iters, long long* cyc) {
  double a[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) a[j] = threadIdx.x * 1e-3 + j;
  double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], b, c);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  double s = 0; for (int j = 0; j < ILP; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void smem_rt(double* out, int iters, long long* cyc) {
  __shared__ double buf[64];
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    buf[threadIdx.x] = v;
    __syncwarp();
    v = buf[(threadIdx.x + 1) & 31] + 1.0;
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = v;
}
__global__ void rcp_chain(double* out, int iters, long long* cyc) {
  double a = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = 1.0 / a + 0.5;
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = a;
}
__device__ __forceinline__ double fast_rcp(double x) {
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); r = fma(r, e, r); return r;
}
__global__ void frcp_chain(double* out, int iters, long long* cyc) {
  double a = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = fast_rcp(a) + 0.5;
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = a;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  long long h; int iters = 2000;
  dep_chain<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency: %.2f cycles\n", (double)h / (iters * 16));
  for (int warps = 1; warps <= 16; warps *= 2) {
    tput<8><<<1, 32 * warps>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA tput ILP8, %2d warps/SM(1 CTA): %.3f warp-DFMA/cycle/SM\n", warps, (double)iters * 4 * 8 * warps / h);
  }
  tput<8><<<148 * 2, 256>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA tput full chip 16 warps/SM: %.3f warp-DFMA/cycle/SM\n", (double)iters * 4 * 8 * 16 / h);
  smem_rt<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("STS->syncwarp->LDS->DADD->syncwarp round trip: %.1f cycles\n", (double)h / iters);
  rcp_chain<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("1.0/x + add chain: %.1f cycles\n", (double)h / iters);
  frcp_chain<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("fast_rcp + add chain: %.1f cycles\n", (double)h / iters);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
