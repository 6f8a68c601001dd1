// C ABI (include/quill_b200.h): context, transcript, device buffers, test and measurement hooks.
// The proving entry points forward to sumcheck.cu / msm.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "blake3.cuh"
#include "ctx.cuh"
#include "ec.cuh"
#include "ff.cuh"
#include "sumcheck.cuh"

namespace qz {
int sumcheck_run(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* tables, int tables_on_device,
                 const qz_expr_node* nodes, size_t n_nodes, const uint8_t* consts, size_t n_consts,
                 const uint8_t* claimed_sum, uint8_t* state, size_t max_coeffs, uint8_t* out_coeffs,
                 uint32_t* out_lens, uint8_t* out_point, uint8_t* out_eval, bool zerocheck, uint8_t* out_z,
                 bool sharded);
int eq_table_device(qz_ctx* ctx, int n, const Fr* z_dev, uint4* out_dev, uint64_t base, uint64_t n_elems);
void comm_destroy(qz_ctx* ctx);
int logup_denominators_run(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* tables, int tables_on_device,
                           const qz_expr_node* nodes_h, size_t n_nodes_h, const qz_expr_node* nodes_m, size_t n_nodes_m,
                           const uint8_t* consts, size_t n_consts, const uint8_t* gamma, void* out, int out_on_device);

// ---- small kernels behind the transcript / test hooks ----------------------------------------------------------------
__global__ void k_draw_fr(uint8_t* state, Fr* out) { *out = tr_draw_fr(state); }
__global__ void k_fr_bytes(const Fr* in, uint8_t* out) { fr_to_le_bytes(*in, out); }
__global__ void k_g1_serialize(const uint8_t* xy, uint8_t* out) {
  // ark-serialize uncompressed G1: x ‖ y canonical LE; 0x80 on the last byte if y > -y; 0x40 + zeros at infinity
  Fq x = fp_load<FqParams>(xy), y = fp_load<FqParams>(xy + 32);
  if (fp_is_zero<FqParams>(x) && fp_is_zero<FqParams>(y)) {
    for (int i = 0; i < 64; i++) out[i] = 0;
    out[63] = 0x40;
    return;
  }
  Fq xc = fp_from_mont<FqParams>(x), yc = fp_from_mont<FqParams>(y), nc = fp_from_mont<FqParams>(fp_neg<FqParams>(y));
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 4; b++) {
      out[4 * i + b] = (uint8_t)(xc.v[i] >> (8 * b));
      out[32 + 4 * i + b] = (uint8_t)(yc.v[i] >> (8 * b));
    }
  bool gt = false;
  for (int i = 7; i >= 0; i--) {
    if (yc.v[i] != nc.v[i]) {
      gt = yc.v[i] > nc.v[i];
      break;
    }
  }
  if (gt) out[63] |= 0x80;
}

template <class P>
__global__ void k_field_op(int op, const uint4* a, const uint4* b, uint4* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<P> x = fp_load<P>(a + 2 * i), y = b ? fp_load<P>(b + 2 * i) : fp_zero<P>(), r;
  switch (op) {
    case 0: r = fp_add<P>(x, y); break;
    case 1: r = fp_sub<P>(x, y); break;
    case 2: r = fp_mul<P>(x, y); break;
    case 3: r = fp_inv<P>(x); break;
    case 4: r = fp_to_mont<P>(x); break;
    case 5: r = fp_from_mont<P>(x); break;
    case 6: {  // deferred reduction: x*y + x*x + y*y from one wide accumulator
      FpWide w;
      wide_zero(w);
      wide_mul_acc<P>(w, x, y);
      wide_mul_acc<P>(w, x, x);
      wide_mul_acc<P>(w, y, y);
      r = wide_reduce<P>(w);
      break;
    }
    case 9: r = fp_sqr<P>(x); break;
    case 10: r = fp_inv_serial<P>(x); break;  // binary extended Euclid (the single-thread critical-path inverse)
    case 8: r = fp_mul2_add<P>(x, y, fp_add<P>(x, y), fp_sub<P>(x, y)); break;  // x*y + (x+y)*(x-y), one reduction
    default: {  // 4096 * (x*y): the accumulator's 17th word in use
      FpWide w;
      wide_zero(w);
      for (int j = 0; j < 4096; j++) wide_mul_acc<P>(w, x, y);
      r = wide_reduce<P>(w);
      break;
    }
  }
  fp_store<P>(out + 2 * i, r);
}

__global__ void k_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Xyzz p = xyzz_from_affine(affine_load(a + 64 * i));
  p = xyzz_add_affine(p, affine_load(b + 64 * i));
  affine_store(out + 64 * i, xyzz_to_affine(p));
}
__global__ void k_g1_mul(const uint8_t* a, const uint4* s, uint8_t* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = fp_from_mont<FrParams>(fp_load<FrParams>(s + 2 * i));
  Xyzz base = xyzz_from_affine(affine_load(a + 64 * i)), acc = xyzz_identity();
  for (int bit = 255; bit >= 0; bit--) {
    acc = xyzz_dbl(acc);
    if ((k.v[bit >> 5] >> (bit & 31)) & 1) acc = xyzz_add(acc, base);
  }
  affine_store(out + 64 * i, xyzz_to_affine(acc));
}

// splitmix64-seeded Fr: 252 random bits (< r), stored as-is (the Montgomery form of a uniformly spread element)
__global__ void k_random_fr(uint4* out, size_t n, uint64_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t w[4];
  uint64_t s = seed + 0x9E3779B97F4A7C15ull * (4 * i + 1);
  for (int j = 0; j < 4; j++) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    w[j] = z ^ (z >> 31);
  }
  w[3] &= 0x0fffffffffffffffull;
  out[2 * i] = make_uint4((uint32_t)w[0], (uint32_t)(w[0] >> 32), (uint32_t)w[1], (uint32_t)(w[1] >> 32));
  out[2 * i + 1] = make_uint4((uint32_t)w[2], (uint32_t)(w[2] >> 32), (uint32_t)w[3], (uint32_t)(w[3] >> 32));
}

// ---- integer-pipe micro-benchmarks -----------------------------------------------------------------------------------------
// variant 0: the field multiplier's inner pattern (IMAD.WIDE.U32 with predicate carry), 4 independent chains/thread
__global__ void __launch_bounds__(256) k_bench_imad_wide(uint32_t* sink, int iters, uint32_t seed) {
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 7, a3 = a0 * 7 + 11, b = seed | 1;
  uint32_t t[4][9];
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < 9; i++) t[c][i] = seed + c * 9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c++)
      mad_chain_even(t[c][0], t[c][1], t[c][2], t[c][3], t[c][4], t[c][5], t[c][6], t[c][7], t[c][8], a0, a1, a2, a3,
                     b + c);
  }
  uint32_t x = 0;
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < 9; i++) x ^= t[c][i];
  if (x == 0x12345678u) sink[0] = x;
}
// variant 1: plain 32-bit IMAD, 8 independent chains/thread
__global__ void __launch_bounds__(256) k_bench_imad32(uint32_t* sink, int iters, uint32_t seed) {
  uint32_t v[8], m = seed | 1, c = seed * 7 + 3;
  for (int i = 0; i < 8; i++) v[i] = seed + threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = v[i] * m + c;
  }
  uint32_t x = 0;
  for (int i = 0; i < 8; i++) x ^= v[i];
  if (x == 0x12345678u) sink[0] = x;
}
template <class P>
__global__ void __launch_bounds__(256) k_bench_fp_mul(uint4* sink, int iters, uint32_t seed) {
  Fp<P> a[2], b;
  for (int i = 0; i < 8; i++) {
    a[0].v[i] = (seed + threadIdx.x * 17 + i) & 0x0fffffffu;
    a[1].v[i] = (seed * 3 + threadIdx.x * 5 + i) & 0x0fffffffu;
    b.v[i] = (seed * 7 + blockIdx.x + i) & 0x0fffffffu;
  }
  for (int it = 0; it < iters; it++) {
    a[0] = fp_mul<P>(a[0], b);
    a[1] = fp_mul<P>(a[1], b);
  }
  Fp<P> s = fp_add<P>(a[0], a[1]);
  if (s.v[0] == 0x12345678u && s.v[7] == 0x9abcdef0u) fp_store<P>(sink, s);
}

}  // namespace qz

using namespace qz;

extern "C" {

const char* qz_status_str(int s) {
  switch (s) {
    case QZ_OK: return "ok";
    case QZ_ERR_INVALID_ARG: return "invalid argument";
    case QZ_ERR_DEGREE: return "Polynomial degree exceeds max degree";
    case QZ_ERR_CUDA: return "CUDA error";
    case QZ_ERR_NCCL: return "NCCL error";
    case QZ_ERR_EXPR: return "malformed or oversized expression";
    case QZ_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
    case QZ_ERR_ALLOC: return "allocation failed";
    default: return "unknown status";
  }
}

int qz_ctx_create(int device, void* stream, qz_ctx** out) {
  if (!out) return QZ_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return QZ_ERR_NO_DEVICE;
  }
  qz_ctx* c = new (std::nothrow) qz_ctx();
  if (!c) return QZ_ERR_ALLOC;
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete c;
    return QZ_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return QZ_ERR_CUDA;
    }
    c->own_stream = true;
  }
  c->pdl = getenv("QZ_NO_PDL") == nullptr;
  cudaEventCreate(&c->ev_call0);
  cudaEventCreate(&c->ev_call1);
  cudaEventCreate(&c->ev_k0);
  cudaEventCreate(&c->ev_k1);
  *out = c;
  return QZ_OK;
}

void qz_ctx_destroy(qz_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  comm_destroy(c);
  for (auto& b : c->blocks) cudaFree(b.p);
  for (auto& kv : c->cache) cudaFree(kv.second);
  c->pool_trim();
  if (c->pinned) cudaFreeHost(c->pinned);
  cudaEventDestroy(c->ev_call0);
  cudaEventDestroy(c->ev_call1);
  cudaEventDestroy(c->ev_k0);
  cudaEventDestroy(c->ev_k1);
  if (c->acc_nonzero_dev) cudaFree(c->acc_nonzero_dev);
  for (auto& pr : c->acc_ring) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  c->destroy_prep_stream();
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* qz_last_error(const qz_ctx* c) { return c ? c->err.c_str() : "null context"; }
uint64_t qz_kernel_launches(const qz_ctx* c) { return c ? c->launches : 0; }
double qz_last_stat(const qz_ctx* c, int which) { return c && which >= 0 && which < 4 ? c->last_stat[which] : -1.0; }
float qz_last_elapsed_ms(qz_ctx* c, int which) { return c && which >= 0 && which < 2 ? c->last_ms[which] : -1.f; }

int qz_msm_accumulate_stats(qz_ctx* c, int reset, double* out_ms, double* out_mixed_additions, uint64_t* out_launches) {
  if (!c) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(c, cudaSetDevice(c->device));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  double total = 0;
  for (size_t i = 0; i < c->acc_ring_used; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->acc_ring[i].first, c->acc_ring[i].second) == cudaSuccess) total += ms;
    else cudaGetLastError();
  }
  if (out_ms) *out_ms = total;
  unsigned long long nz = 0;
  if (c->acc_nonzero_dev) QZ_CUDA(c, cudaMemcpy(&nz, c->acc_nonzero_dev, 8, cudaMemcpyDeviceToHost));
  if (out_mixed_additions) *out_mixed_additions = (double)nz;
  if (out_launches) *out_launches = c->acc_ring_used;
  if (reset) {
    c->acc_ring_used = 0;
    c->acc_ring_adds = 0;
    c->acc_ring_on = reset > 0;  // reset = 1: start (or restart) collecting; reset = -1: stop
    if (c->acc_ring_on && !c->acc_nonzero_dev) QZ_CUDA(c, cudaMalloc(&c->acc_nonzero_dev, 8));
    if (c->acc_nonzero_dev) QZ_CUDA(c, cudaMemset(c->acc_nonzero_dev, 0, 8));
  }
  return QZ_OK;
}

int qz_ctx_sync(qz_ctx* c) {
  if (!c) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return QZ_OK;
}

int qz_dev_alloc(qz_ctx* c, size_t bytes, void** out) {
  if (!c || !out) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(c, cudaSetDevice(c->device));
  *out = c->pool_alloc(bytes);
  if (!*out) return c->fail(QZ_ERR_ALLOC, "cudaMalloc");
  return QZ_OK;
}
int qz_dev_free(qz_ctx* c, void* p) {
  if (!c) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(c, cudaSetDevice(c->device));
  QZ_CUDA(c, c->pool_release(p));
  return QZ_OK;
}
int qz_dev_trim(qz_ctx* c) {
  if (!c) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(c, cudaSetDevice(c->device));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  c->pool_trim();
  return QZ_OK;
}
int qz_dev_upload(qz_ctx* c, void* dev, const void* host, size_t bytes) {
  if (!c || !dev || !host) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(c, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return QZ_OK;
}
int qz_dev_download(qz_ctx* c, void* host, const void* dev, size_t bytes) {
  if (!c || !dev || !host) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(c, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return QZ_OK;
}
int qz_dev_random_fr(qz_ctx* c, void* dev, size_t n, uint64_t seed) {
  if (!c || !dev) return QZ_ERR_INVALID_ARG;
  if (n == 0) return QZ_OK;
  QZ_LAUNCH(c, k_random_fr, (unsigned)((n + 255) / 256), 256, 0, (uint4*)dev, n, seed);
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return QZ_OK;
}

// ---- transcript -----------------------------------------------------------------------------------------------------------------
void qz_transcript_new(const uint8_t* domain, size_t len, uint8_t state[32]) {
  Blake3::hash(domain, len, state, 32);
}
void qz_transcript_append_bytes(uint8_t state[32], const uint8_t* msg, size_t len) {
  std::vector<uint8_t> buf(32 + len);
  memcpy(buf.data(), state, 32);
  if (len) memcpy(buf.data() + 32, msg, len);
  Blake3::hash(buf.data(), buf.size(), state, 32);
}
void qz_transcript_draw_challenge(uint8_t state[32], uint8_t* out, size_t n) {
  uint8_t buf[41];
  memcpy(buf, state, 32);
  memcpy(buf + 32, "challenge", 9);
  Blake3::hash(buf, 41, out, n);
  qz_transcript_append_bytes(state, out, n);
}
int qz_transcript_draw_fr(qz_ctx* c, uint8_t state[32], uint8_t out_fr[32]) {
  if (!c || !state || !out_fr) return QZ_ERR_INVALID_ARG;
  c->arena_reset();
  uint8_t* d = (uint8_t*)c->arena_alloc(64);
  if (!d) return c->fail(QZ_ERR_ALLOC, "scratch");
  QZ_CUDA(c, cudaMemcpyAsync(d, state, 32, cudaMemcpyHostToDevice, c->stream));
  QZ_LAUNCH(c, k_draw_fr, 1, 1, 0, d, (Fr*)(d + 32));
  uint8_t h[64];
  QZ_CUDA(c, cudaMemcpyAsync(h, d, 64, cudaMemcpyDeviceToHost, c->stream));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(state, h, 32);
  memcpy(out_fr, h + 32, 32);
  return QZ_OK;
}
int qz_transcript_append_fr(qz_ctx* c, uint8_t state[32], const uint8_t fr[32]) {
  if (!c || !state || !fr) return QZ_ERR_INVALID_ARG;
  c->arena_reset();
  uint8_t* d = (uint8_t*)c->arena_alloc(64);
  if (!d) return c->fail(QZ_ERR_ALLOC, "scratch");
  QZ_CUDA(c, cudaMemcpyAsync(d, fr, 32, cudaMemcpyHostToDevice, c->stream));
  QZ_LAUNCH(c, k_fr_bytes, 1, 1, 0, (const Fr*)d, d + 32);
  uint8_t h[32];
  QZ_CUDA(c, cudaMemcpyAsync(h, d + 32, 32, cudaMemcpyDeviceToHost, c->stream));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  qz_transcript_append_bytes(state, h, 32);
  return QZ_OK;
}
int qz_g1_serialize(qz_ctx* c, const uint8_t xy[64], uint8_t out[64]) {
  if (!c || !xy || !out) return QZ_ERR_INVALID_ARG;
  c->arena_reset();
  uint8_t* d = (uint8_t*)c->arena_alloc(128);
  if (!d) return c->fail(QZ_ERR_ALLOC, "scratch");
  QZ_CUDA(c, cudaMemcpyAsync(d, xy, 64, cudaMemcpyHostToDevice, c->stream));
  QZ_LAUNCH(c, k_g1_serialize, 1, 1, 0, d, d + 64);
  QZ_CUDA(c, cudaMemcpyAsync(out, d + 64, 64, cudaMemcpyDeviceToHost, c->stream));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return QZ_OK;
}
int qz_transcript_append_g1(qz_ctx* c, uint8_t state[32], const uint8_t xy[64]) {
  uint8_t b[64];
  int rc = qz_g1_serialize(c, xy, b);
  if (rc) return rc;
  qz_transcript_append_bytes(state, b, 64);
  return QZ_OK;
}

// ---- sumcheck ---------------------------------------------------------------------------------------------------------------------
int qz_sumcheck_prove(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* tables, int tables_on_device,
                      const qz_expr_node* nodes, size_t n_nodes, const uint8_t* consts, size_t n_consts,
                      const uint8_t claimed_sum[32], uint8_t state[32], size_t max_coeffs, uint8_t* out_coeffs,
                      uint32_t* out_lens, uint8_t* out_point, uint8_t out_eval[32]) {
  if (!ctx) return QZ_ERR_INVALID_ARG;
  return sumcheck_run(ctx, num_vars, k, tables, tables_on_device, nodes, n_nodes, consts, n_consts, claimed_sum, state,
                      max_coeffs, out_coeffs, out_lens, out_point, out_eval, false, nullptr, false);
}
int qz_sumcheck_prove_sharded(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* table_shards,
                              int tables_on_device, const qz_expr_node* nodes, size_t n_nodes, const uint8_t* consts,
                              size_t n_consts, const uint8_t claimed_sum[32], uint8_t state[32], size_t max_coeffs,
                              uint8_t* out_coeffs, uint32_t* out_lens, uint8_t* out_point, uint8_t out_eval[32]) {
  if (!ctx) return QZ_ERR_INVALID_ARG;
  return sumcheck_run(ctx, num_vars, k, table_shards, tables_on_device, nodes, n_nodes, consts, n_consts, claimed_sum,
                      state, max_coeffs, out_coeffs, out_lens, out_point, out_eval, false, nullptr, true);
}
int qz_zerocheck_prove(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* tables, int tables_on_device,
                       const qz_expr_node* nodes, size_t n_nodes, const uint8_t* consts, size_t n_consts,
                       uint8_t state[32], size_t max_coeffs, uint8_t* out_coeffs, uint32_t* out_lens,
                       uint8_t* out_point, uint8_t out_eval[32], uint8_t* out_z) {
  if (!ctx) return QZ_ERR_INVALID_ARG;
  return sumcheck_run(ctx, num_vars, k, tables, tables_on_device, nodes, n_nodes, consts, n_consts, nullptr, state,
                      max_coeffs, out_coeffs, out_lens, out_point, out_eval, true, out_z, false);
}
int qz_zerocheck_prove_sharded(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* table_shards,
                               int tables_on_device, const qz_expr_node* nodes, size_t n_nodes, const uint8_t* consts,
                               size_t n_consts, uint8_t state[32], size_t max_coeffs, uint8_t* out_coeffs,
                               uint32_t* out_lens, uint8_t* out_point, uint8_t out_eval[32], uint8_t* out_z) {
  if (!ctx) return QZ_ERR_INVALID_ARG;
  return sumcheck_run(ctx, num_vars, k, table_shards, tables_on_device, nodes, n_nodes, consts, n_consts, nullptr, state,
                      max_coeffs, out_coeffs, out_lens, out_point, out_eval, true, out_z, true);
}
int qz_logup_denominators(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* tables, int tables_on_device,
                          const qz_expr_node* nodes_h, size_t n_nodes_h, const qz_expr_node* nodes_m, size_t n_nodes_m,
                          const uint8_t* consts, size_t n_consts, const uint8_t gamma[32], void* out, int out_on_device) {
  if (!ctx) return QZ_ERR_INVALID_ARG;
  return logup_denominators_run(ctx, num_vars, k, tables, tables_on_device, nodes_h, n_nodes_h, nodes_m, n_nodes_m, consts,
                                n_consts, gamma, out, out_on_device);
}
int qz_eq_table(qz_ctx* c, size_t n, const uint8_t* point, void* out, int out_on_device) {
  if (!c || (n && !point) || !out || n >= (size_t)SC_MAX_VARS) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(c, cudaSetDevice(c->device));
  c->arena_reset();
  Fr* z = (Fr*)c->arena_alloc(32 * (n ? n : 1));
  const size_t bytes = (size_t)32 << n;
  uint4* dst = out_on_device ? (uint4*)out : (uint4*)c->arena_alloc(bytes);
  if (!z || !dst) return c->fail(QZ_ERR_ALLOC, "eq table");
  if (n) QZ_CUDA(c, cudaMemcpyAsync(z, point, 32 * n, cudaMemcpyHostToDevice, c->stream));
  int rc = eq_table_device(c, (int)n, z, dst, 0, (uint64_t)1 << n);
  if (rc) return rc;
  if (!out_on_device) QZ_CUDA(c, cudaMemcpyAsync(out, dst, bytes, cudaMemcpyDeviceToHost, c->stream));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return QZ_OK;
}

// ---- test hooks -----------------------------------------------------------------------------------------------------------------------
int qz_test_field_op(qz_ctx* c, int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  if (!c || !a || !out || op < 0 || op > 10 || field < 0 || field > 1) return QZ_ERR_INVALID_ARG;
  if (n == 0) return QZ_OK;
  c->arena_reset();
  uint4* da = (uint4*)c->arena_alloc(32 * n);
  uint4* db = b ? (uint4*)c->arena_alloc(32 * n) : nullptr;
  uint4* dout = (uint4*)c->arena_alloc(32 * n);
  if (!da || !dout || (b && !db)) return c->fail(QZ_ERR_ALLOC, "scratch");
  QZ_CUDA(c, cudaMemcpyAsync(da, a, 32 * n, cudaMemcpyHostToDevice, c->stream));
  if (b) QZ_CUDA(c, cudaMemcpyAsync(db, b, 32 * n, cudaMemcpyHostToDevice, c->stream));
  unsigned grid = (unsigned)((n + 127) / 128);
  if (field == 0)
    QZ_LAUNCH(c, k_field_op<FrParams>, grid, 128, 0, op, da, db, dout, n);
  else
    QZ_LAUNCH(c, k_field_op<FqParams>, grid, 128, 0, op, da, db, dout, n);
  QZ_CUDA(c, cudaMemcpyAsync(out, dout, 32 * n, cudaMemcpyDeviceToHost, c->stream));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return QZ_OK;
}
int qz_test_g1_add(qz_ctx* c, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  if (!c || !a || !b || !out) return QZ_ERR_INVALID_ARG;
  if (n == 0) return QZ_OK;
  c->arena_reset();
  uint8_t *da = (uint8_t*)c->arena_alloc(64 * n), *db = (uint8_t*)c->arena_alloc(64 * n),
          *dd = (uint8_t*)c->arena_alloc(64 * n);
  if (!da || !db || !dd) return c->fail(QZ_ERR_ALLOC, "scratch");
  QZ_CUDA(c, cudaMemcpyAsync(da, a, 64 * n, cudaMemcpyHostToDevice, c->stream));
  QZ_CUDA(c, cudaMemcpyAsync(db, b, 64 * n, cudaMemcpyHostToDevice, c->stream));
  QZ_LAUNCH(c, k_g1_add, (unsigned)((n + 63) / 64), 64, 0, da, db, dd, n);
  QZ_CUDA(c, cudaMemcpyAsync(out, dd, 64 * n, cudaMemcpyDeviceToHost, c->stream));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return QZ_OK;
}
int qz_test_g1_mul(qz_ctx* c, const uint8_t* a, const uint8_t* s, uint8_t* out, size_t n) {
  if (!c || !a || !s || !out) return QZ_ERR_INVALID_ARG;
  if (n == 0) return QZ_OK;
  c->arena_reset();
  uint8_t *da = (uint8_t*)c->arena_alloc(64 * n), *ds = (uint8_t*)c->arena_alloc(32 * n),
          *dd = (uint8_t*)c->arena_alloc(64 * n);
  if (!da || !ds || !dd) return c->fail(QZ_ERR_ALLOC, "scratch");
  QZ_CUDA(c, cudaMemcpyAsync(da, a, 64 * n, cudaMemcpyHostToDevice, c->stream));
  QZ_CUDA(c, cudaMemcpyAsync(ds, s, 32 * n, cudaMemcpyHostToDevice, c->stream));
  QZ_LAUNCH(c, k_g1_mul, (unsigned)((n + 63) / 64), 64, 0, da, (const uint4*)ds, dd, n);
  QZ_CUDA(c, cudaMemcpyAsync(out, dd, 64 * n, cudaMemcpyDeviceToHost, c->stream));
  QZ_CUDA(c, cudaStreamSynchronize(c->stream));
  return QZ_OK;
}

// ---- measurement hooks --------------------------------------------------------------------------------------------------------------------
int qz_bench_imad(qz_ctx* c, int variant, double* out_ops_per_s) {
  if (!c || !out_ops_per_s || variant < 0 || variant > 1) return QZ_ERR_INVALID_ARG;
  c->arena_reset();
  uint32_t* sink = (uint32_t*)c->arena_alloc(256);
  if (!sink) return c->fail(QZ_ERR_ALLOC, "scratch");
  const int iters = 4096, grid = c->sm_count * 8;
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    QZ_CUDA(c, cudaEventRecord(c->ev_k0, c->stream));
    if (variant == 0)
      QZ_LAUNCH(c, k_bench_imad_wide, grid, 256, 0, sink, iters, 12345u + rep);
    else
      QZ_LAUNCH(c, k_bench_imad32, grid, 256, 0, sink, iters, 12345u + rep);
    QZ_CUDA(c, cudaEventRecord(c->ev_k1, c->stream));
    QZ_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1);
    // variant 0: 4 chains x 4 wide multiply-accumulates per iteration; variant 1: 4 x 8 IMADs per iteration
    double ops = (double)grid * 256 * iters * (variant == 0 ? 16.0 : 32.0);
    if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
  }
  *out_ops_per_s = best;
  return QZ_OK;
}
int qz_bench_fp_mul(qz_ctx* c, int field, double* out_muls_per_s) {
  if (!c || !out_muls_per_s || field < 0 || field > 1) return QZ_ERR_INVALID_ARG;
  c->arena_reset();
  uint4* sink = (uint4*)c->arena_alloc(256);
  if (!sink) return c->fail(QZ_ERR_ALLOC, "scratch");
  const int iters = 512, grid = c->sm_count * 8;
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    QZ_CUDA(c, cudaEventRecord(c->ev_k0, c->stream));
    if (field == 0)
      QZ_LAUNCH(c, k_bench_fp_mul<FrParams>, grid, 256, 0, sink, iters, 777u + rep);
    else
      QZ_LAUNCH(c, k_bench_fp_mul<FqParams>, grid, 256, 0, sink, iters, 777u + rep);
    QZ_CUDA(c, cudaEventRecord(c->ev_k1, c->stream));
    QZ_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1);
    double ops = (double)grid * 256 * iters * 2.0;
    if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
  }
  *out_muls_per_s = best;
  return QZ_OK;
}

}  // extern "C"
