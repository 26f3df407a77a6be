// DMMA (mma.sync.m8n8k4.f64) issue/latency probe on sm_100a: W warps per SM (W/4 per SMSP scheduler), each issuing
// NACC independent accumulator chains.  Prints cycles per DMMA per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(double* out, double a, double b, long long* cyc) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = a + i; c[i][1] = b + i; }
  double fa = a + threadIdx.x, fb = b - threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma884(c[i][0], c[i][1], fa, fb);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC> void run(int warps, double* out, long long* dc) {
  k<NACC><<<148, warps * 32>>>(out, 1.0, 0.999, dc);
  cudaDeviceSynchronize();
  k<NACC><<<148, warps * 32>>>(out, 1.0, 0.999, dc);
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  double per_warp = (double)c / (ITERS * NACC);              // cycles per DMMA seen by one warp
  double per_smsp = per_warp / ((warps + 3) / 4);            // cycles per DMMA per scheduler
  printf("{\"warps_per_sm\":%d,\"nacc\":%d,\"cyc_per_dmma_warp\":%.1f,\"cyc_per_dmma_smsp\":%.1f}\n", warps, NACC, per_warp, per_smsp);
}
int main() {
  double* out; long long* dc; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&dc, 8);
  for (int w : {1, 4, 8, 16}) { run<1>(w, out, dc); run<2>(w, out, dc); run<4>(w, out, dc); run<8>(w, out, dc); run<16>(w, out, dc); }
  printf("{\"err\":\"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
}
