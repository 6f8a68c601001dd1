// Single-thread latency of the per-round serial pieces (development tool).
#include <cstdio>
#include <cuda_runtime.h>
#include "sumcheck.cuh"
using namespace qz;

__global__ void klat(uint32_t* state_g, long long* t, Fr* sink) {
  __shared__ uint32_t s_msg[SC_MSG_WORDS];
  if (threadIdx.x != 0) return;
  uint32_t* state = state_g;
  for (int i = 0; i < SC_MSG_WORDS; i++) s_msg[i] = i * 2654435761u;
  for (int i = 10 + 32; i < 48; i++) s_msg[i] = 0;
  long long t0 = clock64();
  tr_absorb_words(state, s_msg, 8 + 32 * 4);
  long long t1 = clock64();
  Fr r = tr_draw_fr_words(state);
  long long t2 = clock64();
  Fr a = r, b = r;
  for (int i = 0; i < 8; i++) a = fp_mul<FrParams>(a, b);
  long long t3 = clock64();
  for (int i = 0; i < 8; i++) a = fp_add<FrParams>(fp_mul<FrParams>(a, b), r);
  long long t4 = clock64();
  uint32_t out[16];
  b3_single_chunk(s_msg, 64, out, nullptr);
  long long t5 = clock64();
  a.v[0] ^= out[3];
  *sink = a;
  t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4;
}

// the same transcript steps on four lanes (tr_*_quad)
__global__ void klat4(uint32_t* state_g, long long* t, Fr* sink) {
  __shared__ uint32_t s_msg[SC_MSG_WORDS];
  if (threadIdx.x >= 4) return;
  for (int i = threadIdx.x; i < SC_MSG_WORDS; i += 4) s_msg[i] = i * 2654435761u;
  __syncwarp(B3_QUAD);
  for (int i = 10 + 32 + threadIdx.x; i < 48; i += 4) s_msg[i] = 0;
  __syncwarp(B3_QUAD);
  long long t0 = clock64();
  tr_absorb_quad(state_g, s_msg, 8 + 32 * 4);
  long long t1 = clock64();
  Fr r = tr_draw_fr_quad(state_g, s_msg);
  long long t2 = clock64();
  if (threadIdx.x == 0) {
    *sink = r;
    t[0] = t1 - t0;
    t[1] = t2 - t1;
  }
}

// ILP probe: do two (four) independent dependent-product chains overlap in one thread?  Cycles per step of 1, 2, 4 chains.
__global__ void kilp(long long* t, Fr* sink, Fr seed) {
  if (threadIdx.x != 0) return;
  Fr a = seed, b = seed, c = seed, d = seed, m = seed;
  a.v[0] ^= 1; b.v[0] ^= 2; c.v[0] ^= 3; d.v[0] ^= 4;
  long long t0 = clock64();
  for (int i = 0; i < 16; i++) a = fp_mul<FrParams>(a, m);
  long long t1 = clock64();
  for (int i = 0; i < 16; i++) { a = fp_mul<FrParams>(a, m); b = fp_mul<FrParams>(b, m); }
  long long t2 = clock64();
  for (int i = 0; i < 16; i++) { a = fp_mul<FrParams>(a, m); b = fp_mul<FrParams>(b, m); c = fp_mul<FrParams>(c, m); d = fp_mul<FrParams>(d, m); }
  long long t3 = clock64();
  *sink = fp_add<FrParams>(fp_add<FrParams>(a, b), fp_add<FrParams>(c, d));
  t[0] = (t1 - t0) / 16; t[1] = (t2 - t1) / 16; t[2] = (t3 - t2) / 16;
}

// dependent-load latency of the three global load flavours the short rounds could use (pointer chase over 4 KiB)
__global__ void kload(const uint32_t* chain, long long* t, uint32_t* sink) {
  if (threadIdx.x != 0) return;
  uint32_t i = 0;
  long long t0 = clock64();
  for (int k = 0; k < 64; k++) i = __ldcv(chain + i);
  long long t1 = clock64();
  for (int k = 0; k < 64; k++) i = __ldcg(chain + i);
  long long t2 = clock64();
  for (int k = 0; k < 64; k++) i = __ldca(chain + i);
  long long t3 = clock64();
  *sink = i;
  t[0] = (t1 - t0) / 64; t[1] = (t2 - t1) / 64; t[2] = (t3 - t2) / 64;
}

int main() {
  uint32_t* st; long long* t; Fr* sink;
  cudaMalloc(&st, 32); cudaMalloc(&t, 64); cudaMalloc(&sink, 32);
  cudaMemset(st, 1, 32);
  for (int rep = 0; rep < 3; rep++) {
    klat<<<1, 32>>>(st, t, sink);
    long long h[5];
    cudaMemcpy(h, t, 40, cudaMemcpyDeviceToHost);
    printf("absorb(3 blocks) %lld cyc | draw_fr %lld cyc | 8 dependent mul %lld cyc | 8 mul+add %lld | 1 compress %lld cyc   %s\n", h[0], h[1], h[2], h[3], h[4],
           cudaGetErrorString(cudaGetLastError()));
  }
  for (int rep = 0; rep < 3; rep++) {
    klat4<<<1, 32>>>(st, t, sink);
    long long h[2];
    cudaMemcpy(h, t, 16, cudaMemcpyDeviceToHost);
    printf("four lanes: absorb(3 blocks) %lld cyc | draw_fr %lld cyc   %s\n", h[0], h[1], cudaGetErrorString(cudaGetLastError()));
  }
  for (int rep = 0; rep < 2; rep++) {
    Fr seed;
    for (int i = 0; i < 8; i++) seed.v[i] = 0x01234567u * (i + 1);
    seed.v[7] &= 0x0fffffffu;
    kilp<<<1, 32>>>(t, sink, seed);
    long long r[3];
    cudaMemcpy(r, t, 24, cudaMemcpyDeviceToHost);
    printf("ILP probe: 1 chain %lld cyc/step | 2 chains %lld | 4 chains %lld   %s\n", r[0], r[1], r[2], cudaGetErrorString(cudaGetLastError()));
  }
  {
    uint32_t h[1024], *d, *sk;
    for (int i = 0; i < 1024; i++) h[i] = (i * 37 + 11) & 1023;  // a permutation of the 1024 words
    cudaMalloc(&d, 4096); cudaMalloc(&sk, 4);
    cudaMemcpy(d, h, 4096, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; rep++) {
      kload<<<1, 32>>>(d, t, sk);
      long long r[3];
      cudaMemcpy(r, t, 24, cudaMemcpyDeviceToHost);
      printf("dependent load: ld.cv %lld cyc | ld.cg %lld cyc | ld.ca %lld cyc\n", r[0], r[1], r[2]);
    }
  }
  return 0;
}
