// Micro-benchmark of Montgomery multiplier variants (development tool; build: nvcc -O3 -std=c++17 -gencode
// arch=compute_100a,code=sm_100a -Itools -o /tmp/mulbench tools/mulbench.cu).
#include <cstdio>
#include <cuda_runtime.h>
#include "ff_variants.cuh"  // fp_mul_inline, fp_sqr_sos (+ the product header ff.cuh)
using namespace qz;

template <class P, int VAR>
__device__ __forceinline__ Fp<P> mulv(const Fp<P>& a, const Fp<P>& b) {
  if (VAR == 0) return fp_mul_inline<P>(a, b);
#ifdef HAVE_V1
  if (VAR == 1) return fp_mul_v1<P>(a, b);
#endif
  return fp_mul_inline<P>(a, b);
}

template <class P, int VAR, int CHAINS>
__global__ void __launch_bounds__(256) kbench(uint4* out, int iters, uint32_t seed) {
  Fp<P> x[CHAINS], y;
  for (int c = 0; c < CHAINS; c++)
    for (int i = 0; i < 8; i++) x[c].v[i] = (seed * (c + 3) + threadIdx.x * 17 + i * 1234567 + blockIdx.x) & 0x0fffffffu;
  for (int i = 0; i < 8; i++) y.v[i] = (seed * 7 + blockIdx.x * 3 + i * 7654321) & 0x0fffffffu;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = mulv<P, VAR>(x[c], y);
  }
  Fp<P> s = x[0];
  for (int c = 1; c < CHAINS; c++) s = fp_add<P>(s, x[c]);
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fp_store<P>(out + 2 * gid, s);
}

template <class P, int VAR, int CHAINS>
double run(const char* name, uint4* out, uint4* href, bool check) {
  int grid = 148 * 8, iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kbench<P, VAR, CHAINS><<<grid, 256>>>(out, 10, 1);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    kbench<P, VAR, CHAINS><<<grid, 256>>>(out, iters, 99);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  double rate = (double)grid * 256 * iters * CHAINS / (best * 1e-3);
  size_t n = (size_t)grid * 256 * 2;
  bool ok = true;
  if (check) {
    uint4* h = (uint4*)malloc(n * 16);
    cudaMemcpy(h, out, n * 16, cudaMemcpyDeviceToHost);
    ok = memcmp(h, href, n * 16) == 0;
    free(h);
  } else {
    cudaMemcpy(href, out, n * 16, cudaMemcpyDeviceToHost);
  }
  cudaError_t err = cudaGetLastError();
  printf("%-28s chains=%d  %8.3f ms  %.3e mul/s  cycles/warp-mul/SMSP=%.0f  %s %s\n", name, CHAINS, best, rate,
         148.0 * 4 * 1.965e9 * 32 / rate, check ? (ok ? "MATCH" : "MISMATCH") : "", err ? cudaGetErrorString(err) : "");
  return rate;
}

int main() {
  size_t n = (size_t)148 * 8 * 256 * 2;
  uint4* out;
  cudaMalloc(&out, n * 16);
  uint4* href = (uint4*)malloc(n * 16);
  run<FqParams, 0, 1>("v0 (shift, MOVs) Fq", out, href, false);
#ifdef HAVE_V1
  run<FqParams, 1, 1>("v1 (even/odd, no shift) Fq", out, href, true);
#endif
  run<FqParams, 0, 2>("v0 Fq", out, href, false);
#ifdef HAVE_V1
  run<FqParams, 1, 2>("v1 Fq", out, href, true);
#endif
  run<FqParams, 0, 4>("v0 Fq", out, href, false);
#ifdef HAVE_V1
  run<FqParams, 1, 4>("v1 Fq", out, href, true);
#endif
  run<FrParams, 0, 2>("v0 Fr", out, href, false);
#ifdef HAVE_V1
  run<FrParams, 1, 2>("v1 Fr", out, href, true);
#endif
  return 0;
}
