// Multilinear PCS opening on the device: MLEvalProof::prove (pcs/src/mlpcs.rs:83-124), with
//   P_r                      = eq table of the evaluation point (mlpcs.rs:52-78; SURVEY 8a5: no FFT needed)
//   evaluation               = <poly, P_r>                                   (mlpcs.rs:91-94)
//   S polynomial             = InnerProductProof::compute_s_polynomial        (pcs/src/ipa.rs:122-157) via an Fr NTT
//   1 commit + 4 KZG opens   = 5 MSMs, all operands stay in HBM between steps (mlpcs.rs:97, 109-113)
// and the transcript schedule of mlpcs.rs:100-107 executed on the device, so the whole opening is one enqueue.
//
// NTT: radix-2, forward decimation-in-frequency (natural -> bit-reversed), inverse decimation-in-time (bit-reversed ->
// natural), so no permutation pass is needed; up to NTT_TILE_LOG stages are fused per pass in shared memory (a pass
// reads and writes every element exactly once, one 32-byte sector each).  The polynomial product is independent of the
// choice of primitive root, so results equal arkworks' FFT-backed `&DensePolynomial * &DensePolynomial`.
#include <algorithm>
#include <cstring>
#include "ctx.cuh"
#include "ec.cuh"
#include "msm.cuh"
#include "sumcheck.cuh"

namespace qz {

int eq_table_device(qz_ctx* ctx, int n, const Fr* z_dev, uint4* out_dev, uint64_t base, uint64_t n_elems);  // sumcheck.cu

constexpr int NTT_TILE_LOG = 10;  // 1024 elements = 32 KB of shared memory per block
constexpr int NTT_THREADS = 256;

QZ_DEV Fr fr_const_root28() {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = FrParams::ROOT28(i);
  return r;
}

// W[j] = w^j for j < count, w = primitive 2^log_m-th root of unity
__global__ void __launch_bounds__(128) ntt_twiddles(int log_m, uint64_t count, Fr* W) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, begin = t * 32;
  if (begin >= count) return;
  Fr w = fr_const_root28();
  for (int i = 28; i > log_m; i--) w = fp_sqr<FrParams>(w);
  Fr cur = fp_one<FrParams>();
  for (int bit = 63 - __clzll(begin | 1); bit >= 0; bit--) {
    cur = fp_sqr<FrParams>(cur);
    if ((begin >> bit) & 1) cur = fp_mul<FrParams>(cur, w);
  }
  const uint64_t end = begin + 32 < count ? begin + 32 : count;
  for (uint64_t j = begin; j < end; j++) {
    W[j] = cur;
    cur = fp_mul<FrParams>(cur, w);
  }
}

// stages [s0, s0 + T) of a size-2^log_m transform.  An element's index reads  hi | t (T bits) | lo (lo_bits bits)  and
// the 2^T elements that share (hi, lo) are closed under those stages: a sub-tile.  A block works on 2^logC sub-tiles at
// once -- the ones whose (hi, lo) differ in the lowest logC bits, i.e. adjacent columns -- so that every pass keeps all
// threads busy whatever T is (a pass of 3 stages used to run 4 threads per block) and a strided pass still moves runs
// of 2^logC consecutive elements.  Shared memory holds element (t, u) of sub-tile u at [t * C + u].
// A 32-byte element read by consecutive threads as two 128-bit words puts threads i and i+4 of a quarter-warp on the
// same banks (ncu: half of all shared wavefronts were conflicts), so the tile keeps the two halves in separate planes.
struct NttTile {
  uint4 lo[1 << NTT_TILE_LOG], hi[1 << NTT_TILE_LOG];
  QZ_DEV Fr get(int i) const {
    const uint4 a = lo[i], b = hi[i];
    Fr r;
    r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
    return r;
  }
  QZ_DEV void set(int i, const Fr& r) {
    lo[i] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    hi[i] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
  }
};
// Twiddles.  Stage s pairs index i with i + m / 2^(s+1) and multiplies the difference by w^((i mod m / 2^(s+1)) 2^s); with
// i = hi | t | lo that exponent is (j_l * stride + lo) << s, one distinct twiddle per butterfly of a strided pass -- a
// 32-byte gather each, which bound those passes (ncu: l1tex 90 % of peak, 0.90 / 0.98 ms against 0.59 ms for the pass
// whose twiddles do not depend on lo).  The exponent splits: w^((j_l * stride) << s) is the same for every sub-tile (a
// table of 2^(T-1) values that lives in L1), and the lo part, w^(lo 2^s), multiplies the odd output of stage s and then
// rides through the later (linear) stages unchanged -- both inputs of a later butterfly went through the same branches
// -- so an element that ends at local index t has collected w^(lo 2^s0 brev_T(t)).  A strided pass is therefore a plain
// local transform plus ONE correction product per element, applied while the tile is stored (forward) or, inverted,
// while it is loaded (inverse): the four-step decomposition, done per pass.
// Where element (t, u) of the tile lives.  The transfer loops walk t with u fixed, i.e. in steps of C 16-byte units per
// plane: with C >= 8 every lane of a quarter-warp lands on the same four banks (ncu r01: 30 M conflicts in 64 M
// wavefronts for the 7-stage pass, C = 8), with C = 4 or 2 every second or fourth.  XOR-ing the column with bits of t
// spreads eight consecutive t over the eight 16-byte bank groups, and the butterflies, which walk u with t fixed, still
// touch one contiguous row (a permutation of it).
QZ_DEV int ntt_slot(int t, int u, int logC) {
  const int C = 1 << logC;
  const int sw = logC >= 3 ? (t & 7) : (t >> (3 - logC)) & (C - 1);
  return t * C + (u ^ sw);
}

template <bool INVERSE>
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass(uint4* data, int log_m, int s0, int T, int logC, const Fr* W) {
  __shared__ NttTile tile;
  const uint64_t m = (uint64_t)1 << log_m, groups = m >> (T + logC);
  const int lo_bits = log_m - s0 - T, C = 1 << logC;
  const int lu = logC < lo_bits ? logC : lo_bits;  // bits of u that fall into lo (contiguous in memory)
  const uint64_t stride = (uint64_t)1 << lo_bits, half_m = m >> 1;
  const int n_elems = 1 << (T + logC);
  for (uint64_t group = blockIdx.x; group < groups; group += gridDim.x) {
    const uint64_t tile0 = group << logC;  // first of the block's sub-tiles; sub-tile id = hi << lo_bits | lo
    // global <-> shared in address order: idx = u_hi | t | u_lo
    for (int idx = threadIdx.x; idx < n_elems; idx += blockDim.x) {
      const int u = (idx & ((1 << lu) - 1)) | ((idx >> (T + lu)) << lu), t = (idx >> lu) & ((1 << T) - 1);
      const uint64_t id = tile0 + u, lo = id & (stride - 1), hi = id >> lo_bits;
      const uint4* src = data + 2 * ((hi << (log_m - s0)) + lo + (uint64_t)t * stride);
      if (INVERSE && lo) {  // undo the correction first: w^-E, E = lo * brev_T(t) << s0 (< m)
        const uint64_t E = (lo * (uint64_t)(__brev((unsigned)t) >> (32 - T))) << s0;
        Fr v = fp_load<FrParams>(src);
        if (E) v = E > half_m ? fp_mul<FrParams>(v, W[m - E]) : fp_neg<FrParams>(fp_mul<FrParams>(v, W[half_m - E]));
        tile.set(ntt_slot(t, u, logC), v);
      } else {
        const int slot = ntt_slot(t, u, logC);
        tile.lo[slot] = src[0];
        tile.hi[slot] = src[1];
      }
    }
    __syncthreads();
    for (int step = 0; step < T; step++) {
      const int ls = INVERSE ? T - 1 - step : step;  // local stage; global stage s = s0 + ls
      const int lhalf = 1 << (T - 1 - ls), s = s0 + ls;
      for (int bb = threadIdx.x; bb < (n_elems >> 1); bb += blockDim.x) {
        const int u = bb & (C - 1), b = bb >> logC;
        const int j_l = b & (lhalf - 1), t0 = ((b - j_l) << 1) + j_l;
        const int i0 = ntt_slot(t0, u, logC), i1 = ntt_slot(t0 + lhalf, u, logC);
        const uint64_t e = ((uint64_t)j_l << lo_bits) << s;  // the part of the twiddle exponent all sub-tiles share, < m/2
        const Fr a = tile.get(i0), v = tile.get(i1);
        if (!INVERSE) {
          tile.set(i0, fp_add<FrParams>(a, v));
          const Fr d = fp_sub<FrParams>(a, v);
          tile.set(i1, e ? fp_mul<FrParams>(d, W[e]) : d);
        } else {
          // w^-e = -w^(m/2 - e)
          const Fr vw = e ? fp_neg<FrParams>(fp_mul<FrParams>(v, W[half_m - e])) : v;
          tile.set(i0, fp_add<FrParams>(a, vw));
          tile.set(i1, fp_sub<FrParams>(a, vw));
        }
      }
      __syncthreads();
    }
    for (int idx = threadIdx.x; idx < n_elems; idx += blockDim.x) {
      const int u = (idx & ((1 << lu) - 1)) | ((idx >> (T + lu)) << lu), t = (idx >> lu) & ((1 << T) - 1);
      const uint64_t id = tile0 + u, lo = id & (stride - 1), hi = id >> lo_bits;
      uint4* dst = data + 2 * ((hi << (log_m - s0)) + lo + (uint64_t)t * stride);
      if (!INVERSE && lo) {  // the correction: w^E, E = lo * brev_T(t) << s0 (< m)
        const uint64_t E = (lo * (uint64_t)(__brev((unsigned)t) >> (32 - T))) << s0;
        Fr v = tile.get(ntt_slot(t, u, logC));
        if (E) v = E < half_m ? fp_mul<FrParams>(v, W[E]) : fp_neg<FrParams>(fp_mul<FrParams>(v, W[E - half_m]));
        fp_store<FrParams>(dst, v);
      } else {
        const int slot = ntt_slot(t, u, logC);
        dst[0] = tile.lo[slot];
        dst[1] = tile.hi[slot];
      }
    }
    __syncthreads();
  }
}

QZ_DEV uint64_t brev_bits(uint64_t x, int bits) { return bits ? (__brevll(x) >> (64 - bits)) : 0; }

// H[k] = w^((L-1) k) (A[k] B[-k] + A[-k] B[k]) on bit-reversed storage (see the derivation in DESIGN.md section 9)
__global__ void __launch_bounds__(256) s_pointwise(const uint4* A, const uint4* B, uint4* H, int log_m, uint64_t l_minus_1,
                                                  const Fr* W) {
  const uint64_t m = (uint64_t)1 << log_m, half_m = m >> 1;
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m) return;
  const uint64_t k = brev_bits(p, log_m), kn = (m - k) & (m - 1), pn = brev_bits(kn, log_m);
  const Fr a = fp_load<FrParams>(A + 2 * p), b = fp_load<FrParams>(B + 2 * p);
  const Fr an = fp_load<FrParams>(A + 2 * pn), bn = fp_load<FrParams>(B + 2 * pn);
  Fr t = fp_add<FrParams>(fp_mul<FrParams>(a, bn), fp_mul<FrParams>(an, b));
  const uint64_t e = (l_minus_1 * k) & (m - 1);
  if (e) {
    t = e < half_m ? fp_mul<FrParams>(t, W[e]) : fp_neg<FrParams>(fp_mul<FrParams>(t, W[e - half_m]));
  }
  fp_store<FrParams>(H + 2 * p, t);
}
// S[k] = h[L + k] / m, k < L - 1.  1/m = (1/2)^log_m is computed once (s_minv), not by every thread.
__global__ void s_minv(int log_m, Fr* out) {
  Fr inv2, minv = fp_one<FrParams>();
#pragma unroll
  for (int i = 0; i < 8; i++) inv2.v[i] = FrParams::INV2(i);
  for (int i = 0; i < log_m; i++) minv = fp_mul<FrParams>(minv, inv2);
  *out = minv;
}
__global__ void __launch_bounds__(256) s_extract(const uint4* h, uint64_t L, const Fr* minv_p, uint4* S) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k + 1 >= L) return;
  const Fr minv = *minv_p;
  fp_store<FrParams>(S + 2 * k, fp_mul<FrParams>(fp_load<FrParams>(h + 2 * (L + k)), minv));
}
// sum_i a[i] * b[i], i < n: one partial per block, then `inner_product_final`
__global__ void __launch_bounds__(256) inner_product_partial(const uint4* a, const uint4* b, uint64_t n, Fr* partials) {
  __shared__ Fr s_warp[32];
  Fr acc = fp_zero<FrParams>();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    acc = fp_add<FrParams>(acc, fp_mul<FrParams>(fp_load<FrParams>(a + 2 * i), fp_load<FrParams>(b + 2 * i)));
  block_sum_to(acc, s_warp, &partials[blockIdx.x]);
}
__global__ void __launch_bounds__(256) inner_product_final(const Fr* partials, int n_parts, Fr* out) {
  __shared__ Fr s_warp[32];
  Fr acc = fp_zero<FrParams>();
  for (int i = threadIdx.x; i < n_parts; i += blockDim.x) acc = fp_add<FrParams>(acc, partials[i]);
  block_sum_to(acc, s_warp, out);
}
// index of the last non-zero element + 1 (DensePolynomial's trimmed length)
__global__ void __launch_bounds__(256) fr_trimmed_len(const uint4* a, uint64_t n, unsigned long long* out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long best = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (!fp_is_zero<FrParams>(fp_load<FrParams>(a + 2 * i))) best = i + 1;
  if (best) atomicMax(out, best);
}

// ark-serialize uncompressed G1 (x ‖ y canonical LE, 0x80 if y > -y, 0x40 + zeros at infinity)
QZ_DEV void g1_serialize_dev(const uint8_t* xy, uint8_t* out) {
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
  for (int i = 7; i >= 0; i--)
    if (yc.v[i] != nc.v[i]) {
      gt = yc.v[i] > nc.v[i];
      break;
    }
  if (gt) out[63] |= 0x80;
}
// mlpcs.rs:100-107: absorb &[F] point (length-prefixed), evaluation, S commitment; squeeze r; r_inv
__global__ void mlpcs_transcript(uint8_t* state, const Fr* point, int n, const Fr* evaluation, const uint8_t* s_comm,
                                 Fr* r_out) {
  uint8_t buf[8 + 32 * SC_MAX_VARS];
  for (int i = 0; i < 8; i++) buf[i] = (uint8_t)((uint64_t)n >> (8 * i));
  for (int i = 0; i < n; i++) fr_to_le_bytes(point[i], buf + 8 + 32 * i);
  tr_absorb(state, buf, 8 + 32 * n);
  fr_to_le_bytes(*evaluation, buf);
  tr_absorb(state, buf, 32);
  g1_serialize_dev(s_comm, buf);
  tr_absorb(state, buf, 64);
  const Fr r = tr_draw_fr(state);
  r_out[0] = r;
  r_out[1] = fp_inv_serial<FrParams>(r);  // r = 0 has probability 2^-254; the reference would panic on unwrap (:107)
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
static int get_twiddles(qz_ctx* ctx, int log_m, Fr** out) {
  // cached per context (same map as the interpolation matrices, keys offset by 1000), one table per size and never
  // evicted: a prover alternates between a few sizes (HyperPlonk: 2^21, 2^23, 2^24 within one proof) and rebuilding the
  // table at every switch cost a cudaFree (device-wide synchronisation) + cudaMalloc + the kernel.  All sizes together
  // are at most twice the largest table (256 MiB at 2^24).
  auto it = ctx->cache.find(1000 + log_m);
  if (it != ctx->cache.end()) {
    *out = (Fr*)it->second;
    return QZ_OK;
  }
  const uint64_t count = log_m ? ((uint64_t)1 << (log_m - 1)) : 1;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, 32 * count);
  if (e != cudaSuccess) return ctx->fail(QZ_ERR_ALLOC, "twiddle table", e);
  QZ_LAUNCH(ctx, ntt_twiddles, (unsigned)((count + 32 * 128 - 1) / (32 * 128)), 128, 0, log_m, count, (Fr*)p);
  ctx->cache[1000 + log_m] = p;
  *out = (Fr*)p;
  return QZ_OK;
}

int ntt_device(qz_ctx* ctx, uint4* data, int log_m, bool inverse) {
  if (log_m == 0) return QZ_OK;
  QzRange nvtx_call(inverse ? "qz:ntt:inverse" : "qz:ntt:forward");
  if (log_m > 28) return ctx->fail(QZ_ERR_INVALID_ARG, "NTT size exceeds the two-adicity of Fr (2^28)");
  Fr* W = nullptr;
  int rc = get_twiddles(ctx, log_m, &W);
  if (rc) return rc;
  // stage groups [s0, s0 + T), as few as fit the tile and of (almost) equal size: forward in increasing order, inverse
  // in decreasing order
  int starts[32], sizes[32];
  const int ng = (log_m + NTT_TILE_LOG - 1) / NTT_TILE_LOG;
  for (int g = 0, s0 = 0; g < ng; g++) {
    starts[g] = s0;
    sizes[g] = log_m / ng + (g < log_m % ng ? 1 : 0);
    s0 += sizes[g];
  }
  for (int gi = 0; gi < ng; gi++) {
    const int g = inverse ? ng - 1 - gi : gi;
    const int logC = std::min(NTT_TILE_LOG, log_m) - sizes[g];
    const uint64_t groups = ((uint64_t)1 << log_m) >> (sizes[g] + logC);
    const unsigned grid = (unsigned)std::min<uint64_t>(groups, (uint64_t)ctx->sm_count * 8);
    if (inverse)
      QZ_LAUNCH(ctx, ntt_pass<true>, grid, NTT_THREADS, 0, data, log_m, starts[g], sizes[g], logC, W);
    else
      QZ_LAUNCH(ctx, ntt_pass<false>, grid, NTT_THREADS, 0, data, log_m, starts[g], sizes[g], logC, W);
  }
  return QZ_OK;
}

// S polynomial of (f, n) and (g, m) on the device: writes max(n, m) - 1 coefficients (not trimmed) to S
int s_polynomial_device(qz_ctx* ctx, const uint4* f, size_t n, const uint4* g, size_t m_len, uint4* S) {
  QzRange nvtx_call("qz:mlpcs:s-polynomial");
  const uint64_t L = std::max(n, m_len);
  if (L < 2) return QZ_OK;
  int log_m = 1;
  while (((uint64_t)1 << log_m) < 2 * L - 1) log_m++;
  if (log_m > 28) return ctx->fail(QZ_ERR_INVALID_ARG, "S polynomial needs an NTT beyond 2^28");
  const uint64_t M = (uint64_t)1 << log_m;
  auto mark = ctx->arena_mark();
  uint4* A = (uint4*)ctx->arena_alloc(32 * M);
  uint4* B = (uint4*)ctx->arena_alloc(32 * M);
  uint4* H = (uint4*)ctx->arena_alloc(32 * M);
  Fr* minv = (Fr*)ctx->arena_alloc(32);
  if (!A || !B || !H || !minv) return ctx->fail(QZ_ERR_ALLOC, "NTT buffers");
  cudaStream_t st = ctx->stream;
  QZ_CUDA(ctx, cudaMemsetAsync(A, 0, 32 * M, st));
  QZ_CUDA(ctx, cudaMemsetAsync(B, 0, 32 * M, st));
  if (n) QZ_CUDA(ctx, cudaMemcpyAsync(A, f, 32 * n, cudaMemcpyDeviceToDevice, st));
  if (m_len) QZ_CUDA(ctx, cudaMemcpyAsync(B, g, 32 * m_len, cudaMemcpyDeviceToDevice, st));
  int rc = ntt_device(ctx, A, log_m, false);
  if (rc) return rc;
  rc = ntt_device(ctx, B, log_m, false);
  if (rc) return rc;
  Fr* W = nullptr;
  rc = get_twiddles(ctx, log_m, &W);
  if (rc) return rc;
  QZ_LAUNCH(ctx, s_pointwise, (unsigned)((M + 255) / 256), 256, 0, A, B, H, log_m, L - 1, W);
  rc = ntt_device(ctx, H, log_m, true);
  if (rc) return rc;
  QZ_LAUNCH(ctx, s_minv, 1, 1, 0, log_m, minv);
  QZ_LAUNCH(ctx, s_extract, (unsigned)((L + 255) / 256), 256, 0, H, L, (const Fr*)minv, S);
  ctx->arena_release(mark);
  return QZ_OK;
}

int inner_product_device(qz_ctx* ctx, const uint4* a, const uint4* b, size_t n, Fr* out) {
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 4));
  auto mark = ctx->arena_mark();
  Fr* partials = (Fr*)ctx->arena_alloc(32 * (size_t)grid);
  if (!partials) return ctx->fail(QZ_ERR_ALLOC, "inner product partials");
  QZ_LAUNCH(ctx, inner_product_partial, grid, 256, 0, a, b, (uint64_t)n, partials);
  QZ_LAUNCH(ctx, inner_product_final, 1, 256, 0, partials, grid, out);
  ctx->arena_release(mark);
  return QZ_OK;
}

}  // namespace qz

using namespace qz;

extern "C" {

int qz_compute_s_polynomial(qz_ctx* ctx, const uint8_t* p1, size_t n1, const uint8_t* p2, size_t n2, uint8_t* out) {
  if (!ctx || (n1 && !p1) || (n2 && !p2) || !out) return QZ_ERR_INVALID_ARG;
  const size_t L = std::max(n1, n2);
  if (L < 2) return QZ_OK;
  QZ_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->arena_reset();
  cudaStream_t st = ctx->stream;
  uint4* f = (uint4*)ctx->arena_alloc(32 * std::max<size_t>(n1, 1));
  uint4* g = (uint4*)ctx->arena_alloc(32 * std::max<size_t>(n2, 1));
  uint4* S = (uint4*)ctx->arena_alloc(32 * (L - 1));
  if (!f || !g || !S) return ctx->fail(QZ_ERR_ALLOC, "S polynomial");
  if (n1) QZ_CUDA(ctx, cudaMemcpyAsync(f, p1, 32 * n1, cudaMemcpyHostToDevice, st));
  if (n2) QZ_CUDA(ctx, cudaMemcpyAsync(g, p2, 32 * n2, cudaMemcpyHostToDevice, st));
  int rc = s_polynomial_device(ctx, f, n1, g, n2, S);
  if (rc) return rc;
  QZ_CUDA(ctx, cudaMemcpyAsync(out, S, 32 * (L - 1), cudaMemcpyDeviceToHost, st));
  QZ_CUDA(ctx, cudaStreamSynchronize(st));
  return QZ_OK;
}

// MLEvalProof::prove (mlpcs.rs:83-124) in the two halves its Fiat-Shamir challenge separates.
// begin: P_r, evaluation, S, commit(S) (:86-97).  Everything is enqueued on the context's stream; `d_s` must hold
// max(n, 2^n_point) elements, *s_commit_len receives the length commit(S)/open(S) work on.
static int mlpcs_begin_device(qz_ctx* ctx, const qz_srs* srs, const uint4* pdev, size_t n, const uint8_t* point,
                              size_t n_point, Fr* d_point, uint4* d_pr, uint4* d_s, uint8_t* d_eval, uint8_t* d_scomm,
                              uint64_t* s_commit_len) {
  cudaStream_t st = ctx->stream;
  QzRange nvtx_call("qz:mlpcs:begin (P_r, evaluation, S, commit S)");
  // trimmed length of P_r from the point alone: coefficient j vanishes iff some bit i of j is set with r_i = 0 or clear
  // with r_i = 1, so the last non-zero coefficient is j = sum_{r_i != 0} 2^i (always non-zero)
  uint64_t pr_len = 1;
  for (size_t i = 0; i < n_point; i++) {
    bool zero = true;
    for (int b = 0; b < 32; b++) zero &= point[32 * i + b] == 0;
    if (!zero) pr_len += (uint64_t)1 << i;
  }
  const uint64_t pr_full = (uint64_t)1 << n_point;
  const uint64_t L = std::max<uint64_t>(n, pr_len);
  const uint64_t s_len = L >= 1 ? L - 1 : 0;  // untrimmed S length
  if (n_point) QZ_CUDA(ctx, cudaMemcpyAsync(d_point, point, 32 * n_point, cudaMemcpyHostToDevice, st));
  int rc = eq_table_device(ctx, (int)n_point, d_point, d_pr, 0, pr_full);  // P_r coefficients (mlpcs.rs:68-78)
  if (rc) return rc;
  const uint64_t ip_n = std::min<uint64_t>(n, pr_full);  // zip stops at the shorter slice (:92)
  if (ip_n) {
    rc = inner_product_device(ctx, pdev, d_pr, ip_n, (Fr*)d_eval);
    if (rc) return rc;
  }
  if (s_len) {
    rc = s_polynomial_device(ctx, pdev, n, d_pr, pr_len, d_s);  // ipa.rs:122-157
    if (rc) return rc;
  }
  // commit(S) asserts on the TRIMMED length (kzg.rs:62-65); only look it up when the untrimmed one does not fit
  *s_commit_len = s_len;
  if (s_len > srs->n) {
    unsigned long long* d_trim = (unsigned long long*)ctx->arena_alloc(8);
    if (!d_trim) return ctx->fail(QZ_ERR_ALLOC, "mlpcs state");
    QZ_CUDA(ctx, cudaMemsetAsync(d_trim, 0, 8, st));
    QZ_LAUNCH(ctx, fr_trimmed_len, (unsigned)std::min<uint64_t>((s_len + 255) / 256, 4096), 256, 0, d_s, s_len, d_trim);
    unsigned long long t = 0;
    QZ_CUDA(ctx, cudaMemcpyAsync(&t, d_trim, 8, cudaMemcpyDeviceToHost, st));
    QZ_CUDA(ctx, cudaStreamSynchronize(st));
    if (t > srs->n) return ctx->fail(QZ_ERR_DEGREE, "Polynomial degree exceeds max degree");
    *s_commit_len = t;
  }
  if (n > srs->n + 1) return ctx->fail(QZ_ERR_DEGREE, "Polynomial degree exceeds max degree");  // open(poly): quotient
  return msm_device(ctx, srs, d_s, *s_commit_len, nullptr, d_scomm);  // mlpcs.rs:97
}
// finish: poly_opening, poly_opening_inv, s_opening, s_opening_inv at d_r[0] = r and d_r[1] = 1/r (mlpcs.rs:109-113)
static int mlpcs_finish_device(qz_ctx* ctx, const qz_srs* srs, const uint4* pdev, size_t n, const uint4* d_s,
                               uint64_t s_commit_len, const Fr* d_r, uint8_t* d_open) {
  QzRange nvtx_call("qz:mlpcs:finish (4 KZG openings)");
  for (int i = 0; i < 4; i++) {
    uint8_t* o = d_open + 128 * i;
    const Fr* x = d_r + (i & 1);
    QZ_CUDA(ctx, cudaMemcpyAsync(o, x, 32, cudaMemcpyDeviceToDevice, ctx->stream));
    const int rc = i < 2 ? kzg_open_device(ctx, srs, pdev, n, x, (Fr*)(o + 32), o + 64)
                         : kzg_open_device(ctx, srs, d_s, s_commit_len, x, (Fr*)(o + 32), o + 64);
    if (rc) return rc;
  }
  return QZ_OK;
}
static int mlpcs_poly_on_device(qz_ctx* ctx, const void* poly, size_t n, int on_device, const uint4** pdev) {
  *pdev = (const uint4*)poly;
  if (!on_device && n) {
    void* p = ctx->arena_alloc(32 * n);
    if (!p) return ctx->fail(QZ_ERR_ALLOC, "poly");
    QZ_CUDA(ctx, cudaMemcpyAsync(p, poly, 32 * n, cudaMemcpyHostToDevice, ctx->stream));
    *pdev = (const uint4*)p;
  }
  return QZ_OK;
}
__global__ void fr_with_inverse(Fr* r) { r[1] = fp_inv<FrParams>(r[0]); }

int qz_mlpcs_open(qz_ctx* ctx, const qz_srs* srs, const void* poly, size_t n, int on_device, const uint8_t* point,
                  size_t n_point, uint8_t state[32], uint8_t out_evaluation[32], uint8_t out_s_comm[64],
                  uint8_t out_openings[512]) {
  if (!ctx || !srs || (n && !poly) || (n_point && !point) || !state || !out_evaluation || !out_s_comm || !out_openings)
    return QZ_ERR_INVALID_ARG;
  if (n_point >= (size_t)SC_MAX_VARS || n_point > 27) return ctx->fail(QZ_ERR_INVALID_ARG, "too many variables");
  QZ_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->arena_reset();
  cudaStream_t st = ctx->stream;
  QZ_CUDA(ctx, cudaEventRecord(ctx->ev_call0, st));
  QZ_CUDA(ctx, cudaEventRecord(ctx->ev_k0, st));
  QZ_CUDA(ctx, cudaEventRecord(ctx->ev_k1, st));
  const uint64_t pr_full = (uint64_t)1 << n_point;
  uint8_t* res = (uint8_t*)ctx->arena_alloc(32 + 64 + 4 * 128 + 32 + 64);  // eval ‖ s_comm ‖ 4 x (x ‖ y ‖ proof) ‖ state ‖ r, r_inv
  Fr* d_point = (Fr*)ctx->arena_alloc(32 * std::max<size_t>(n_point, 1));
  uint4* d_pr = (uint4*)ctx->arena_alloc(32 * pr_full);
  uint4* d_s = (uint4*)ctx->arena_alloc(32 * std::max<uint64_t>(std::max<uint64_t>(n, pr_full), 1));
  if (!res || !d_point || !d_pr || !d_s) return ctx->fail(QZ_ERR_ALLOC, "mlpcs state");
  uint8_t* d_eval = res;
  uint8_t* d_scomm = res + 32;
  uint8_t* d_open = res + 96;
  uint8_t* d_state = res + 96 + 512;
  Fr* d_r = (Fr*)(res + 96 + 512 + 32);
  QZ_CUDA(ctx, cudaMemsetAsync(res, 0, 96 + 512 + 32 + 64, st));
  QZ_CUDA(ctx, cudaMemcpyAsync(d_state, state, 32, cudaMemcpyHostToDevice, st));
  const uint4* pdev = nullptr;
  int rc = mlpcs_poly_on_device(ctx, poly, n, on_device, &pdev);
  if (rc) return rc;
  uint64_t s_commit_len = 0;
  rc = mlpcs_begin_device(ctx, srs, pdev, n, point, n_point, d_point, d_pr, d_s, d_eval, d_scomm, &s_commit_len);
  if (rc) return rc;
  QZ_LAUNCH(ctx, mlpcs_transcript, 1, 1, 0, d_state, d_point, (int)n_point, (const Fr*)d_eval, d_scomm, d_r);
  rc = mlpcs_finish_device(ctx, srs, pdev, n, d_s, s_commit_len, d_r, d_open);
  if (rc) return rc;
  uint8_t* pin = (uint8_t*)ctx->pinned_buf(96 + 512 + 32);
  if (!pin) return ctx->fail(QZ_ERR_ALLOC, "pinned");
  QZ_CUDA(ctx, cudaMemcpyAsync(pin, res, 96 + 512 + 32, cudaMemcpyDeviceToHost, st));
  QZ_CUDA(ctx, cudaEventRecord(ctx->ev_call1, st));
  QZ_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(out_evaluation, pin, 32);
  memcpy(out_s_comm, pin + 32, 64);
  memcpy(out_openings, pin + 96, 512);
  memcpy(state, pin + 96 + 512, 32);
  cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev_call0, ctx->ev_call1);
  cudaEventElapsedTime(&ctx->last_ms[1], ctx->ev_k0, ctx->ev_k1);
  return QZ_OK;
}

int qz_mlpcs_open_begin(qz_ctx* ctx, const qz_srs* srs, const void* poly, size_t n, int on_device, const uint8_t* point,
                        size_t n_point, uint8_t out_evaluation[32], uint8_t out_s_comm[64], void** out_s_dev,
                        size_t* out_s_len) {
  if (!ctx || !srs || (n && !poly) || (n_point && !point) || !out_evaluation || !out_s_comm || !out_s_dev || !out_s_len)
    return QZ_ERR_INVALID_ARG;
  if (n_point >= (size_t)SC_MAX_VARS || n_point > 27) return ctx->fail(QZ_ERR_INVALID_ARG, "too many variables");
  QZ_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->arena_reset();
  cudaStream_t st = ctx->stream;
  const uint64_t pr_full = (uint64_t)1 << n_point;
  uint8_t* res = (uint8_t*)ctx->arena_alloc(96);
  Fr* d_point = (Fr*)ctx->arena_alloc(32 * std::max<size_t>(n_point, 1));
  uint4* d_pr = (uint4*)ctx->arena_alloc(32 * pr_full);
  if (!res || !d_point || !d_pr) return ctx->fail(QZ_ERR_ALLOC, "mlpcs state");
  // outlives the call: owned by the caller (qz_dev_free)
  void* s_dev = ctx->pool_alloc(32 * std::max<uint64_t>(std::max<uint64_t>(n, pr_full), 1));
  if (!s_dev) return ctx->fail(QZ_ERR_ALLOC, "S polynomial");
  QZ_CUDA(ctx, cudaMemsetAsync(res, 0, 96, st));
  const uint4* pdev = nullptr;
  int rc = mlpcs_poly_on_device(ctx, poly, n, on_device, &pdev);
  uint64_t s_commit_len = 0;
  if (!rc) rc = mlpcs_begin_device(ctx, srs, pdev, n, point, n_point, d_point, d_pr, (uint4*)s_dev, res, res + 32, &s_commit_len);
  uint8_t* pin = rc ? nullptr : (uint8_t*)ctx->pinned_buf(96);
  if (!rc && !pin) rc = ctx->fail(QZ_ERR_ALLOC, "pinned");
  if (rc) {
    cudaStreamSynchronize(st);
    ctx->pool_release(s_dev);
    return rc;
  }
  QZ_CUDA(ctx, cudaMemcpyAsync(pin, res, 96, cudaMemcpyDeviceToHost, st));
  QZ_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(out_evaluation, pin, 32);
  memcpy(out_s_comm, pin + 32, 64);
  *out_s_dev = s_dev;
  *out_s_len = (size_t)s_commit_len;
  return QZ_OK;
}

int qz_mlpcs_open_finish(qz_ctx* ctx, const qz_srs* srs, const void* poly, size_t n, int on_device, const void* s_dev,
                         size_t s_len, const uint8_t r[32], uint8_t out_openings[512]) {
  if (!ctx || !srs || (n && !poly) || (s_len && !s_dev) || !r || !out_openings) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->arena_reset();
  cudaStream_t st = ctx->stream;
  uint8_t* d_open = (uint8_t*)ctx->arena_alloc(512);
  Fr* d_r = (Fr*)ctx->arena_alloc(64);
  if (!d_open || !d_r) return ctx->fail(QZ_ERR_ALLOC, "mlpcs state");
  QZ_CUDA(ctx, cudaMemsetAsync(d_open, 0, 512, st));
  QZ_CUDA(ctx, cudaMemcpyAsync(d_r, r, 32, cudaMemcpyHostToDevice, st));
  QZ_LAUNCH(ctx, fr_with_inverse, 1, 1, 0, d_r);  // mlpcs.rs:107
  const uint4* pdev = nullptr;
  int rc = mlpcs_poly_on_device(ctx, poly, n, on_device, &pdev);
  if (rc) return rc;
  rc = mlpcs_finish_device(ctx, srs, pdev, n, (const uint4*)s_dev, s_len, d_r, d_open);
  if (rc) return rc;
  uint8_t* pin = (uint8_t*)ctx->pinned_buf(512);
  if (!pin) return ctx->fail(QZ_ERR_ALLOC, "pinned");
  QZ_CUDA(ctx, cudaMemcpyAsync(pin, d_open, 512, cudaMemcpyDeviceToHost, st));
  QZ_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(out_openings, pin, 512);
  return QZ_OK;
}

}  // extern "C"
