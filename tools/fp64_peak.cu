// FP64 pipe yardstick for sm_100a: DFMA vs mma.sync f64 shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Prints one JSON line per variant with TFLOP/s (2 flops per FMA).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_dfma(double* out, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = a + i + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fma(x[i], b, a);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_mma884(double* out, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = a + i; c[i][1] = b + i; }
  double fa = a + threadIdx.x, fb = b - threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma884(c[i][0], c[i][1], fa, fb);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_mma1684(double* out, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = a + i + j;
  double fa[2] = {a + threadIdx.x, a - threadIdx.x}; double fb = b - threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma1684(c[i], fa, fb);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_mma1688(double* out, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = a + i + j;
  double fa[4] = {a + threadIdx.x, a - threadIdx.x, a * 2, a * 3}; double fb[2] = {b - threadIdx.x, b};
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma1688(c[i], fa, fb);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_mma16816(double* out, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = a + i + j;
  double fa[8], fb[4];
  for (int j = 0; j < 8; j++) fa[j] = a + threadIdx.x * j;
  for (int j = 0; j < 4; j++) fb[j] = b - threadIdx.x * j;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma16816(c[i], fa, fb);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: DMMA + DFMA interleaved, to see whether the pipes are shared
__global__ void k_mixed(double* out, double a, double b) {
  double c[8][2]; double x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = a + i; c[i][1] = b + i; x[i] = a * i + threadIdx.x; }
  double fa = a + threadIdx.x, fb = b - threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { mma884(c[i][0], c[i][1], fa, fb); x[i] = fma(x[i], b, a); }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int threads = 512;
  for (int bpsm = 1; bpsm <= 4; bpsm *= 2) {
    int blocks = sms * bpsm;
    double warps = (double)blocks * threads / 32;
    float ms;
    ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, 1.0, 0.999); });
    printf("{\"variant\":\"dfma\",\"blocks_per_sm\":%d,\"tflops\":%.2f}\n", bpsm, warps * 32 * 16.0 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_mma884<8><<<blocks, threads>>>(out, 1.0, 0.999); });
    printf("{\"variant\":\"mma.m8n8k4 x8acc\",\"blocks_per_sm\":%d,\"tflops\":%.2f}\n", bpsm, warps * 8 * 256.0 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_mma884<16><<<blocks, threads>>>(out, 1.0, 0.999); });
    printf("{\"variant\":\"mma.m8n8k4 x16acc\",\"blocks_per_sm\":%d,\"tflops\":%.2f}\n", bpsm, warps * 16 * 256.0 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_mma1684<8><<<blocks, threads>>>(out, 1.0, 0.999); });
    printf("{\"variant\":\"mma.m16n8k4 x8acc\",\"blocks_per_sm\":%d,\"tflops\":%.2f}\n", bpsm, warps * 8 * 512.0 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_mma1688<8><<<blocks, threads>>>(out, 1.0, 0.999); });
    printf("{\"variant\":\"mma.m16n8k8 x8acc\",\"blocks_per_sm\":%d,\"tflops\":%.2f}\n", bpsm, warps * 8 * 1024.0 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_mma16816<8><<<blocks, threads>>>(out, 1.0, 0.999); });
    printf("{\"variant\":\"mma.m16n8k16 x8acc\",\"blocks_per_sm\":%d,\"tflops\":%.2f}\n", bpsm, warps * 8 * 2048.0 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_mixed<<<blocks, threads>>>(out, 1.0, 0.999); });
    printf("{\"variant\":\"mixed mma884+dfma\",\"blocks_per_sm\":%d,\"tflops\":%.2f}\n", bpsm, warps * 8 * (256.0 + 32.0) * ITERS * 2 / ms / 1e9);
  }
  printf("{\"sms\":%d,\"clock_khz\":%d,\"err\":\"%s\"}\n", sms, p.clockRate, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
