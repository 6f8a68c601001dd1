// Device-side building blocks of the sumcheck prover: transcript, expression interpreter, block reductions.
// Follows hyperplonk/src/piops/sumcheck.rs:28-114 and transcript/src/transcript.rs:14-75 of the reference.
#pragma once
#include "../../include/quill_b200.h"
#include "blake3.cuh"
#include "ff.cuh"

#ifndef SC_TRACE
#define SC_TRACE(id)  // measurement hook, active only in the -DQZ_SC_TRACE build of sumcheck.cu
#endif

namespace qz {

constexpr int SC_MAX_DEG = QZ_MAX_ROUND_COEFFS - 1;
constexpr int SC_MAX_COEFFS = QZ_MAX_ROUND_COEFFS;
constexpr int SC_MAX_K = 16;      // tables referenced by one expression
constexpr int SC_MAX_OPS = 512;   // postfix program length
constexpr int SC_MAX_STACK = 16;  // interpreter stack depth
constexpr int SC_MAX_VARS = 40;
constexpr int SC_TAIL_LOG = 11;  // from 2^11 elements on one block runs a round alone (sumcheck.cu sc_mid)

enum : uint32_t { SC_OP_IN = 0, SC_OP_CONST = 1, SC_OP_ADD = 2, SC_OP_MUL = 3 };

// VirtualPolyExpr (virtual_polynomial.rs:9-18) compiled to postfix; op = (code << 16) | arg
struct ScProgram {
  uint32_t n_ops, degree, k, n_consts;
  uint32_t ops[SC_MAX_OPS];
};

struct ScTables {
  const uint4* in[SC_MAX_K];
  uint4* out[SC_MAX_K];
};

// proof state that lives on the device for the whole proof
struct ScHead {
  alignas(16) uint8_t tstate[32];  // Transcript.state
  Fr r;                // challenge of the last closed round (pending fold)
  Fr evaluation;       // EvaluationClaim.evaluation
  // eq-factored zero-check (sumcheck.cu "zero-check fast path"): P_j = prod_{i<j} eq(r_i, z_i) after round j-1 closed,
  // and P_{j-1}, the value it had one round earlier (the hand-over to sc_mid needs it)
  Fr zc_prefix, zc_prefix_prev;
  // running claim for the rounds that do not sum X = 1 (sumcheck.cu ProdAcc): s_j(r_j) of the last closed round -- on
  // the eq-factored zero-check path t_j(r_j), the factor of s_j that the weighted sums are taken of
  Fr claim;
  uint32_t peer_fault;     // a peer-mailbox wait timed out during this proof (comm.cuh): the results are void
  uint32_t zc_degenerate;  // a zero-check challenge z_j was 0: no 1 / z_j, the proof is redone without that shortcut
};

// a field element of the proof head read past L1: the previous round may have been closed by another block or kernel
QZ_DEV Fr fr_ld_cv(const Fr* p) {
  const uint4 a = __ldcv(reinterpret_cast<const uint4*>(p)), b = __ldcv(reinterpret_cast<const uint4*>(p) + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// ---- serialization / transcript (device) -------------------------------------------------------------------------
QZ_DEV void fr_to_le_bytes(const Fr& mont, uint8_t* out) {  // ark-serialize: 32 B little-endian canonical
  Fr c = fp_from_mont<FrParams>(mont);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[4 * i + 0] = (uint8_t)(c.v[i]);
    out[4 * i + 1] = (uint8_t)(c.v[i] >> 8);
    out[4 * i + 2] = (uint8_t)(c.v[i] >> 16);
    out[4 * i + 3] = (uint8_t)(c.v[i] >> 24);
  }
}

// Transcript::append_bytes (transcript.rs:26-32): state <- blake3(state ‖ msg)
static __device__ __noinline__ void tr_absorb(uint8_t* state, const uint8_t* msg, uint32_t n) {
  uint8_t buf[32 + 8 + 32 * SC_MAX_VARS];  // largest message: a length-prefixed vector of SC_MAX_VARS elements
  for (int i = 0; i < 32; i++) buf[i] = state[i];
  for (uint32_t i = 0; i < n; i++) buf[32 + i] = msg[i];
  uint8_t out[32];
  Blake3::hash(buf, 32 + n, out, 32);
  for (int i = 0; i < 32; i++) state[i] = out[i];
}

// PrimeField::from_le_bytes_mod_order on 48 bytes -> Montgomery Fr:  x = lo + hi*2^256,  x*R = lo*R + hi*R^2
QZ_DEV Fr fr_from_48_le_bytes(const uint8_t* c) {
  Fr lo, hi, r2, r3;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    lo.v[i] = (uint32_t)c[4 * i] | ((uint32_t)c[4 * i + 1] << 8) | ((uint32_t)c[4 * i + 2] << 16) |
              ((uint32_t)c[4 * i + 3] << 24);
    hi.v[i] = 0;
    r2.v[i] = FrParams::R2(i);
    r3.v[i] = FrParams::R3(i);
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
    hi.v[i] = (uint32_t)c[32 + 4 * i] | ((uint32_t)c[32 + 4 * i + 1] << 8) | ((uint32_t)c[32 + 4 * i + 2] << 16) |
              ((uint32_t)c[32 + 4 * i + 3] << 24);
  // first operand of fp_mul must be < r; the second may be any 256-bit value
  return fp_add<FrParams>(fp_mul<FrParams>(r2, lo), fp_mul<FrParams>(r3, hi));
}

// Transcript::draw_field_element::<Fr> (transcript.rs:49-75): squeeze 48 bytes, re-absorb them, reduce mod r
static __device__ __noinline__ Fr tr_draw_fr(uint8_t* state) {
  uint8_t buf[32 + 9];
  for (int i = 0; i < 32; i++) buf[i] = state[i];
  const char tag[10] = "challenge";
  for (int i = 0; i < 9; i++) buf[32 + i] = (uint8_t)tag[i];
  uint8_t c[48];
  Blake3::hash(buf, 41, c, 48);
  tr_absorb(state, c, 48);
  return fr_from_48_le_bytes(c);
}

// ---- word-oriented transcript for the per-round critical path -------------------------------------------------------------
// The byte-oriented routines above cost ~25 us per call on one device thread (local-memory byte traffic).  Every message
// on the sumcheck path is a whole number of 32-bit words and, with the 32-byte state in front, fits one blake3 chunk
// (<= 1024 bytes) for round polynomials of up to 30 coefficients, so the hot path hashes straight from a word buffer:
// buf[0..8) = state, buf[8..) = message, zero padded to a multiple of 16 words.
// The finalize step runs once per launch on ONE thread, so it executes cold: a fully unrolled compression (7 rounds x 8 G,
// ~1000 straight-line instructions per call site) made instruction fetch the dominant cost (ncu r01: 77k cycles for
// 13.7k instructions).  This copy is a loop over the rounds (the message is permuted in registers between rounds), compiled once.
#define QZ_G(a, b, c, d, mx, my)      \
  a = a + b + (mx);                   \
  d = __funnelshift_r(d ^ a, d ^ a, 16); \
  c = c + d;                          \
  b = __funnelshift_r(b ^ c, b ^ c, 12); \
  a = a + b + (my);                   \
  d = __funnelshift_r(d ^ a, d ^ a, 8);  \
  c = c + d;                          \
  b = __funnelshift_r(b ^ c, b ^ c, 7);
// cv <- first 8 words of compress(cv, m, counter 0, len, flags); out12 (optional) <- first 12 words of the output block
static __device__ __noinline__ void b3_compress_loop(uint32_t* cv, const uint32_t* m, uint32_t block_len, uint32_t flags,
                                                     uint32_t* out12) {
  uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
  uint32_t s8 = Blake3::iv(0), s9 = Blake3::iv(1), s10 = Blake3::iv(2), s11 = Blake3::iv(3);
  uint32_t s12 = 0, s13 = 0, s14 = block_len, s15 = flags;
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = m[i];
#pragma unroll 1
  for (int r = 0; r < 7; r++) {  // one copy of the round in the instruction stream; the message stays in registers
    QZ_G(s0, s4, s8, s12, w[0], w[1])
    QZ_G(s1, s5, s9, s13, w[2], w[3])
    QZ_G(s2, s6, s10, s14, w[4], w[5])
    QZ_G(s3, s7, s11, s15, w[6], w[7])
    QZ_G(s0, s5, s10, s15, w[8], w[9])
    QZ_G(s1, s6, s11, s12, w[10], w[11])
    QZ_G(s2, s7, s8, s13, w[12], w[13])
    QZ_G(s3, s4, s9, s14, w[14], w[15])
    // message permutation between rounds
    const uint32_t t0 = w[2], t1 = w[6], t2 = w[3], t3 = w[10], t4 = w[7], t5 = w[0], t6 = w[4], t7 = w[13];
    const uint32_t t8 = w[1], t9 = w[11], t10 = w[12], t11 = w[5], t12 = w[9], t13 = w[14], t14 = w[15], t15 = w[8];
    w[0] = t0; w[1] = t1; w[2] = t2; w[3] = t3; w[4] = t4; w[5] = t5; w[6] = t6; w[7] = t7;
    w[8] = t8; w[9] = t9; w[10] = t10; w[11] = t11; w[12] = t12; w[13] = t13; w[14] = t14; w[15] = t15;
  }
  if (out12) {
    out12[8] = s8 ^ cv[0];
    out12[9] = s9 ^ cv[1];
    out12[10] = s10 ^ cv[2];
    out12[11] = s11 ^ cv[3];
  }
  cv[0] = s0 ^ s8;  cv[1] = s1 ^ s9;  cv[2] = s2 ^ s10;  cv[3] = s3 ^ s11;
  cv[4] = s4 ^ s12; cv[5] = s5 ^ s13; cv[6] = s6 ^ s14;  cv[7] = s7 ^ s15;
  if (out12)
    for (int i = 0; i < 8; i++) out12[i] = cv[i];
}
#undef QZ_G
// single-chunk hash of `total_bytes` bytes held as words in buf (zero padded to a multiple of 16 words);
// out8 <- digest words; out12 (optional) <- first 48 bytes of the XOF output
static __device__ __noinline__ void b3_single_chunk(const uint32_t* buf, uint32_t total_bytes, uint32_t* out8, uint32_t* out12) {
  uint32_t cv[8];
  for (int i = 0; i < 8; i++) cv[i] = Blake3::iv(i);
  const uint32_t nblocks = total_bytes == 0 ? 1 : (total_bytes + 63) / 64;
#pragma unroll 1
  for (uint32_t b = 0; b < nblocks; b++) {
    const bool last = b + 1 == nblocks;
    const uint32_t flags = (b == 0 ? Blake3::F_CHUNK_START : 0u) | (last ? (Blake3::F_CHUNK_END | Blake3::F_ROOT) : 0u);
    b3_compress_loop(cv, buf + 16 * b, last ? total_bytes - 64 * b : 64u, flags, last ? out12 : nullptr);
  }
  if (out8)
    for (int i = 0; i < 8; i++) out8[i] = cv[i];
}
// state <- blake3(state ‖ msg): msg = buf[8 .. 8 + msg_bytes/4), zero padded by the caller up to a 64-byte boundary
static __device__ __noinline__ void tr_absorb_words(uint32_t* state, uint32_t* buf, uint32_t msg_bytes) {
  for (int i = 0; i < 8; i++) buf[i] = state[i];
  if (32 + msg_bytes <= 1024) {
    b3_single_chunk(buf, 32 + msg_bytes, state, nullptr);
  } else {  // multi-chunk message: general tree hash over the same bytes (little-endian words == bytes)
    uint8_t o[32];
    Blake3::hash(reinterpret_cast<const uint8_t*>(buf), 32 + msg_bytes, o, 32);
    for (int i = 0; i < 8; i++)
      state[i] = (uint32_t)o[4 * i] | ((uint32_t)o[4 * i + 1] << 8) | ((uint32_t)o[4 * i + 2] << 16) | ((uint32_t)o[4 * i + 3] << 24);
  }
}
// draw_field_element::<Fr> (transcript.rs:49-75) on words: XOF-48 of state ‖ "challenge", re-absorb, reduce mod r
static __device__ __noinline__ Fr tr_draw_fr_words(uint32_t* state) {
  uint32_t buf[32], c[12];
  for (int i = 0; i < 32; i++) buf[i] = 0;
  for (int i = 0; i < 8; i++) buf[i] = state[i];
  buf[8] = 0x6c616863u;   // "chal"
  buf[9] = 0x676e656cu;   // "leng"
  buf[10] = 0x00000065u;  // "e"
  b3_single_chunk(buf, 41, nullptr, c);  // first 48 bytes of the root output block
  for (int i = 0; i < 12; i++) buf[8 + i] = c[i];
  b3_single_chunk(buf, 80, state, nullptr);
  Fr lo, hi, r2, r3;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    lo.v[i] = c[i];
    hi.v[i] = i < 4 ? c[8 + i] : 0u;
    r2.v[i] = FrParams::R2(i);
    r3.v[i] = FrParams::R3(i);
  }
  return fp_add<FrParams>(fp_mul<FrParams>(r2, lo), fp_mul<FrParams>(r3, hi));
}

// ---- the same transcript on FOUR lanes -----------------------------------------------------------------------------------
// One lane runs a compression in ~2500 cycles: ~1100 instructions from a warp that can issue every other cycle at best.
// blake3's state is a 4x4 word matrix whose columns (then diagonals) are mixed independently, so lane i of a quad keeps
// column i (a, b, c, d) = (v[i], v[4+i], v[8+i], v[12+i]): the column step is one G per lane, three shuffles rotate b, c, d
// into the diagonals, one more G, three shuffles rotate back -- 7 x (2 G + 6 SHFL) ~ 300 instructions per lane.  The
// chaining value stays spread over the lanes (cva = cv[i], cvb = cv[4+i]) from block to block.
// Called by lanes 0..3 of a warp (mask 0xF); `m` points to 16 message words all four lanes can read (shared memory).
// message schedule: nibble j of B3_SCHED[r] = index of the message word used at position j of round r
__device__ constexpr uint64_t B3_SCHED[7] = {0xfedcba9876543210ull, 0x8fe95cb1d407a362ull, 0x18fb0956e72dca43ull,
    0x61852b04fd3e9c7aull, 0x461035278eafb9dcull, 0x7462a03d1fc85be9ull,
    0xd743c2ae689105fbull};
QZ_DEV uint32_t b3_sched(int r, int j) { return (uint32_t)(B3_SCHED[r] >> (4 * j)) & 15u; }
constexpr unsigned B3_QUAD = 0xFu;
#define QZ_G4(mx, my)                       \
  a = a + b + (mx);                         \
  d = __funnelshift_r(d ^ a, d ^ a, 16);    \
  c = c + d;                                \
  b = __funnelshift_r(b ^ c, b ^ c, 12);    \
  a = a + b + (my);                         \
  d = __funnelshift_r(d ^ a, d ^ a, 8);     \
  c = c + d;                                \
  b = __funnelshift_r(b ^ c, b ^ c, 7);
// (cva, cvb) <- words (lane, 4 + lane) of compress(cv, m, counter 0, block_len, flags); returns word 8 + lane of the
// output block (the extended output the challenge squeeze reads)
static __device__ __noinline__ uint32_t b3_compress_quad(uint32_t& cva, uint32_t& cvb, const uint32_t* m, uint32_t block_len,
                                                         uint32_t flags) {
  const int lane = threadIdx.x & 3, l1 = (lane + 1) & 3, l2 = (lane + 2) & 3, l3 = (lane + 3) & 3;
  uint32_t w[7][4];
#pragma unroll
  for (int r = 0; r < 7; r++) {
    w[r][0] = m[b3_sched(r, 2 * lane)];
    w[r][1] = m[b3_sched(r, 2 * lane + 1)];
    w[r][2] = m[b3_sched(r, 8 + 2 * lane)];
    w[r][3] = m[b3_sched(r, 9 + 2 * lane)];
  }
  uint32_t a = cva, b = cvb, c = Blake3::iv(lane), d = lane == 2 ? block_len : lane == 3 ? flags : 0u;
#pragma unroll
  for (int r = 0; r < 7; r++) {
    QZ_G4(w[r][0], w[r][1])
    b = __shfl_sync(B3_QUAD, b, l1, 4);
    c = __shfl_sync(B3_QUAD, c, l2, 4);
    d = __shfl_sync(B3_QUAD, d, l3, 4);
    QZ_G4(w[r][2], w[r][3])
    b = __shfl_sync(B3_QUAD, b, l3, 4);
    c = __shfl_sync(B3_QUAD, c, l2, 4);
    d = __shfl_sync(B3_QUAD, d, l1, 4);
  }
  const uint32_t x = c ^ cva;
  cva = a ^ c;
  cvb = b ^ d;
  return x;
}
#undef QZ_G4
// single-chunk hash of buf[0 .. total_bytes/4) (zero padded to a multiple of 16 words, <= 1024 bytes) on a quad:
// lane i ends with digest words (i, 4 + i) in (cva, cvb) and, in *x, word 8 + i of the root output block
QZ_DEV void b3_single_chunk_quad(const uint32_t* buf, uint32_t total_bytes, uint32_t& cva, uint32_t& cvb, uint32_t* x) {
  const int lane = threadIdx.x & 3;
  cva = Blake3::iv(lane);
  cvb = Blake3::iv(4 + lane);
  const uint32_t nblocks = total_bytes == 0 ? 1 : (total_bytes + 63) / 64;
  uint32_t xo = 0;
#pragma unroll 1
  for (uint32_t blk = 0; blk < nblocks; blk++) {
    const bool last = blk + 1 == nblocks;
    const uint32_t flags = (blk == 0 ? Blake3::F_CHUNK_START : 0u) | (last ? (Blake3::F_CHUNK_END | Blake3::F_ROOT) : 0u);
    xo = b3_compress_quad(cva, cvb, buf + 16 * blk, last ? total_bytes - 64 * blk : 64u, flags);
  }
  if (x) *x = xo;
}
// Transcript::append_bytes on a quad: buf (shared) holds the message at [8, 8 + msg_bytes/4), zero padded to a block
// boundary; state (8 words, any memory the quad can read and write) <- blake3(state ‖ msg).  32 + msg_bytes <= 1024.
QZ_DEV void tr_absorb_quad(uint32_t* state, uint32_t* buf, uint32_t msg_bytes) {
  const int lane = threadIdx.x & 3;
  buf[lane] = __ldcv(state + lane);  // past L1: another block (sc_mid) or kernel may have written the state
  buf[4 + lane] = __ldcv(state + 4 + lane);
  __syncwarp(B3_QUAD);
  uint32_t cva, cvb;
  b3_single_chunk_quad(buf, 32 + msg_bytes, cva, cvb, nullptr);
  state[lane] = cva;
  state[4 + lane] = cvb;
  __syncwarp(B3_QUAD);
}
// draw_field_element::<Fr> (transcript.rs:49-75) on a quad; buf: 32 words of shared memory.  Every lane returns r.
QZ_DEV Fr tr_draw_fr_quad(uint32_t* state, uint32_t* buf) {
  const int lane = threadIdx.x & 3;
  for (int i = lane; i < 32; i += 4) buf[i] = i < 8 ? __ldcv(state + i) : 0u;
  __syncwarp(B3_QUAD);
  if (lane == 0) {
    buf[8] = 0x6c616863u;   // "chal"
    buf[9] = 0x676e656cu;   // "leng"
    buf[10] = 0x00000065u;  // "e"
  }
  __syncwarp(B3_QUAD);
  uint32_t cva, cvb, x;
  b3_single_chunk_quad(buf, 41, cva, cvb, &x);  // 48 bytes of extended output: words lane, 4 + lane, 8 + lane
  __syncwarp(B3_QUAD);
  buf[8 + lane] = cva;
  buf[12 + lane] = cvb;
  buf[16 + lane] = x;
  __syncwarp(B3_QUAD);
  uint32_t na, nb;
  b3_single_chunk_quad(buf, 80, na, nb, nullptr);  // re-absorb the 48 challenge bytes (transcript.rs:60)
  state[lane] = na;
  state[4 + lane] = nb;
  // x = lo + hi 2^256 -> x R = lo R^2 / R + hi R^3 / R: lane 0 takes the first product, lane 1 the second
  Fr op, cst;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    op.v[i] = (lane & 1) ? (i < 4 ? buf[16 + i] : 0u) : buf[8 + i];
    cst.v[i] = (lane & 1) ? FrParams::R3(i) : FrParams::R2(i);
  }
  __syncwarp(B3_QUAD);
  Fr part = fp_mul<FrParams>(cst, op), other;
#pragma unroll
  for (int i = 0; i < 8; i++) other.v[i] = __shfl_xor_sync(B3_QUAD, part.v[i], 1, 4);
  return fp_add<FrParams>(part, other);
}

// ---- reductions ------------------------------------------------------------------------------------------------------
QZ_DEV Fr warp_sum(Fr v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Fr o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.v[i] = __shfl_down_sync(0xffffffffu, v.v[i], off);
    v = fp_add<FrParams>(v, o);
  }
  return v;
}
// sum over the block; *dst written by thread 0.  s_warp: 32 Fr of shared memory.  Ends with a barrier.
QZ_DEV void block_sum_to(Fr v, Fr* s_warp, Fr* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < nwarps ? s_warp[lane] : fp_zero<FrParams>();
    v = warp_sum(v);
    if (lane == 0) *dst = v;
  }
  __syncthreads();
}

// sum n values per thread over the block with ONE pair of barriers: warps reduce by shuffles, lane 0 parks the warp's
// n sums in s_part[warp * n + x], then thread x adds the per-warp sums.  dst[x] written by thread x (x < n).
// s_part: (blockDim.x / 32) * n Fr of shared memory.  Ends with a barrier.
QZ_DEV void block_sum_many(const Fr* vals, int n, Fr* s_part, Fr* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  for (int x = 0; x < n; x++) {
    const Fr v = warp_sum(vals[x]);
    if (lane == 0) s_part[warp * n + x] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    Fr v = s_part[threadIdx.x];
    for (int w = 1; w < nwarps; w++) v = fp_add<FrParams>(v, s_part[w * n + threadIdx.x]);
    dst[threadIdx.x] = v;
  }
  __syncthreads();
}

// ---- expression interpreter (virtual_polynomial.rs:22-37 evaluated on scalars) ---------------------------------------
static __device__ __noinline__ Fr sc_eval_program(const uint32_t* ops, uint32_t n_ops, const Fr* consts, const Fr* vals) {
  Fr stack[SC_MAX_STACK];
  int sp = 0;
  for (uint32_t pc = 0; pc < n_ops; pc++) {
    const uint32_t op = ops[pc], code = op >> 16, arg = op & 0xffffu;
    if (code == SC_OP_IN) {
      stack[sp++] = vals[arg];
    } else if (code == SC_OP_CONST) {
      stack[sp++] = consts[arg];
    } else {
      Fr b = stack[--sp], a = stack[sp - 1];
      stack[sp - 1] = code == SC_OP_ADD ? fp_add<FrParams>(a, b) : fp_mul<FrParams>(a, b);
    }
  }
  return stack[0];
}

// Close a round (sumcheck.rs:67-78): evaluations at X = 0..d (shared s_evals) -> monomial coefficients, trimmed
// length, absorb `len ‖ coeffs`, squeeze the challenge.  Called by every thread of the block; blockDim.x > d.
// s_msg: SC_MSG_WORDS words of shared memory: [0..8) state, [8..10) u64 length, then 8 words per coefficient.
constexpr int SC_MSG_WORDS = ((10 + 8 * SC_MAX_COEFFS + 15) / 16) * 16;
constexpr int SC_PROD_SLOTS = 256;
// zc_z (optional, eq-factored zero-check): the evaluations are those of t_j(X) = sum_x' E_{j+1}(x') prod_t g_t(X, x'), of
// degree d, and the round polynomial is s_j(X) = P_j * eq(X, z_j) * t_j(X), of degree d + 1, with z_j = *zc_z and
// P_j = head->zc_prefix: the d + 1 coefficients are multiplied by the linear factor a + b X, a = P_j (1 - z_j),
// b = P_j (2 z_j - 1), before they are trimmed and absorbed, and P_{j+1} = P_j eq(r_j, z_j) once r_j is drawn.
// Row t of the table of shifted multiples of a challenge for the next pass's folds (ff.cuh fp_mul_fixed):
// C_t = r * 2^e_t mod p, e_t = fold_consts_exponent(t), canonical.  With rM = r R: fp_mul(rM, x) = r x for a raw x.
QZ_DEV void fold_table_row(const Fr& rM, int t, uint32_t* foldc) {
  const int e = fold_consts_exponent(t);
  Fr x = fp_zero<FrParams>(), c;
  if (e < 256) {
    x.v[e >> 5] = 1u << (e & 31);
    c = fp_mul<FrParams>(rM, x);
  } else if (e == 256) {
    c = rM;  // r * 2^256 mod p is the Montgomery form itself
  } else {   // e == 288: (r * 2^32) * 2^256
    Fr r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.v[i] = FrParams::R2(i);
    x.v[1] = 1u;
    c = fp_mul<FrParams>(fp_mul<FrParams>(rM, x), r2);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) foldc[8 * t + i] = c.v[i];
}

// Restore the value at X = 1 that a SKIP1 round did not sum.  s_evals holds d sums compactly (X = 0, 2, .., d) and is
// expanded in place to X = 0..d.  Plain rounds: s_j(1) = s_{j-1}(r_{j-1}) - s_j(0).  Eq-factored zero-check rounds:
// t_{j-1}(r_{j-1}) = sum_x E_j(x) prod_t g_t = (1 - z_j) t_j(0) + z_j t_j(1), so t_j(1) = (claim - (1 - z_j) t_j(0)) / z_j.
// Called by every thread of the block; ends with a barrier.
QZ_DEV void sc_expand_evals(const ScHead* head, int d, Fr* s_evals, const Fr* zc_z, const Fr* zc_zinv) {
  Fr mine = fp_zero<FrParams>();
  const int t = threadIdx.x;
  if (t >= 2 && t <= d) mine = s_evals[t - 1];
  if (t == 1) {
    const Fr e0 = s_evals[0], claim = fr_ld_cv(&head->claim);
    if (zc_z) {
      const Fr z = *zc_z, one = fp_one<FrParams>();
      mine = fp_mul<FrParams>(fp_sub<FrParams>(claim, fp_mul<FrParams>(fp_sub<FrParams>(one, z), e0)), *zc_zinv);
    } else {
      mine = fp_sub<FrParams>(claim, e0);
    }
  }
  __syncthreads();
  if (t >= 1 && t <= d) s_evals[t] = mine;
  __syncthreads();
}

// Samples of a cubic at X = 0, 1, -1 and infinity (its leading coefficient), in that order in s_evals[0..4), to its
// coefficients, in place: c0 = s(0), c3 = s(inf), c2 = (s(1) + s(-1)) / 2 - c0, c1 = (s(1) - s(-1)) / 2 - c3.
// Called by every thread of the block; ends with a barrier.
QZ_DEV void sc_toom3_to_coeffs(Fr* s_evals) {
  const int t = threadIdx.x;
  Fr mine = fp_zero<FrParams>();
  if (t == 1 || t == 2) {
    const Fr s0 = s_evals[0], s1 = s_evals[1], sm = s_evals[2], sinf = s_evals[3];
    Fr inv2;
#pragma unroll
    for (int i = 0; i < 8; i++) inv2.v[i] = FrParams::INV2(i);
    mine = t == 2 ? fp_sub<FrParams>(fp_mul<FrParams>(inv2, fp_add<FrParams>(s1, sm)), s0)
                  : fp_sub<FrParams>(fp_mul<FrParams>(inv2, fp_sub<FrParams>(s1, sm)), sinf);
  }
  __syncthreads();
  if (t == 1 || t == 2) s_evals[t] = mine;
  __syncthreads();
}

// want_claim: leave s_j(r_j) (t_j(r_j) on the eq-factored path) in head->claim for a following SKIP1 round
// release_flag (sc_mid): set to release_value as soon as the challenge is in head->r -- the other blocks of the grid
// start folding with it while this block finishes the bookkeeping.  That bookkeeping (the running claim: d dependent
// products; the zero-check prefix) is done by lane 0 of the LAST warp, which a short round leaves without work, so
// the first warps of this block are not held up either.
QZ_DEV void sc_round_close(ScHead* head, const Fr* vinv, int d, const Fr* s_evals, Fr* s_coef, uint32_t* s_msg, Fr* s_prod,
                           Fr* out_coeffs_row, uint32_t* out_len, Fr* out_point_slot, int max_coeffs,
                           const Fr* zc_z = nullptr, bool want_claim = false, uint32_t* foldc = nullptr,
                           unsigned int* release_flag = nullptr, unsigned int release_value = 0) {
  __shared__ Fr s_chal;
  const int t = threadIdx.x;
  const int n1 = d + 1;
  const int n_out = zc_z ? d + 2 : d + 1;  // coefficients of the round polynomial before trimming
  // coefficient i = sum_j vinv[i][j] * e[j]: the (d+1)^2 products are independent, so when they fit the block each
  // thread does one (s_prod: SC_PROD_SLOTS Fr of shared memory)
  // (vinv null: s_evals already holds the coefficients, see sc_toom3_to_coeffs)
  const bool wide = vinv && n1 * n1 <= (int)blockDim.x && n1 * n1 <= SC_PROD_SLOTS;
  if (wide) {
    if (t < n1 * n1) s_prod[t] = fp_mul<FrParams>(vinv[t], s_evals[t % n1]);
    __syncthreads();
  }
  Fr acc = fp_zero<FrParams>();
  if (t <= d) {
    if (!vinv) {
      acc = s_evals[t];
    } else if (wide) {
      for (int j = 0; j <= d; j++) acc = fp_add<FrParams>(acc, s_prod[t * n1 + j]);
    } else {
#pragma unroll 1
      for (int j = 0; j <= d; j++) acc = fp_add<FrParams>(acc, fp_mul<FrParams>(vinv[t * (d + 1) + j], s_evals[j]));
    }
    s_coef[t] = acc;
  }
  SC_TRACE(10);
  Fr zc_P = fp_zero<FrParams>();
  if (zc_z) {  // (c_0 + c_1 X + ...)(a + b X): c'_t = a c_t + b c_{t-1}
    __syncthreads();
    Fr lower = fp_zero<FrParams>();
    if (t >= 1 && t <= d + 1) lower = s_coef[t - 1];
    if (want_claim && t <= d) s_prod[t] = acc;  // t_j's coefficients, for the running claim (s_prod is free again)
    __syncthreads();
    zc_P = fr_ld_cv(&head->zc_prefix);
    if (t < n_out) {
      const Fr z = *zc_z, one = fp_one<FrParams>();
      const Fr a = fp_mul<FrParams>(zc_P, fp_sub<FrParams>(one, z));
      const Fr b = fp_mul<FrParams>(zc_P, fp_sub<FrParams>(fp_dbl<FrParams>(z), one));
      acc = fp_add<FrParams>(fp_mul<FrParams>(a, acc), fp_mul<FrParams>(b, lower));  // acc = c_t (zero for t = d + 1)
      s_coef[t] = acc;
    }
  }
  if (t < n_out) {
    out_coeffs_row[t] = acc;
    const Fr can = fp_from_mont<FrParams>(acc);  // ark-serialize: 32 B little-endian canonical
#pragma unroll
    for (int i = 0; i < 8; i++) s_msg[10 + 8 * t + i] = can.v[i];
  } else if (t < max_coeffs) {
    out_coeffs_row[t] = fp_zero<FrParams>();
  }
  __syncthreads();
  SC_TRACE(11);
  if (t < 4) {  // the transcript runs on lanes 0..3 of warp 0 (tr_*_quad)
    if (t == 0) {
      int len = n_out;
      while (len > 0 && fp_is_zero<FrParams>(s_coef[len - 1])) len--;  // DensePolynomial trims trailing zeros
      *out_len = (uint32_t)len;
      s_msg[8] = (uint32_t)len;  // u64 LE length prefix
      s_msg[9] = 0;
      for (int i = 10 + 8 * len; i < ((10 + 8 * len + 15) / 16) * 16; i++) s_msg[i] = 0;  // pad the last block
    }
    __syncwarp(B3_QUAD);
    const uint32_t len = s_msg[8];
    uint32_t* state = reinterpret_cast<uint32_t*>(head->tstate);
    if (32 + 8 + 32 * len <= 1024) {
      tr_absorb_quad(state, s_msg, 8 + 32 * len);  // :73
    } else {  // 31 coefficients or more: state ‖ message spans two blake3 chunks, which the quad routine does not hash
      if (t == 0) tr_absorb_words(state, s_msg, 8 + 32 * len);
      __syncwarp(B3_QUAD);
    }
    SC_TRACE(12);
    const Fr r = tr_draw_fr_quad(state, s_msg);   // :77
    SC_TRACE(13);
    if (t == 0) {
      head->r = r;
      *out_point_slot = r;
      s_chal = r;
    }
  }
  __syncthreads();
  if (t == 0 && release_flag) {  // every write of this round (tables, state, challenge) before the flag
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(release_flag), "r"(release_value) : "memory");
  }
  if (t == (int)blockDim.x - 32) {
    const Fr r = s_chal;
    if (want_claim) {  // Horner at the challenge
      const Fr* c = zc_z ? s_prod : s_coef;
      Fr v = c[d];
      for (int i = d - 1; i >= 0; i--) v = fp_add<FrParams>(fp_mul<FrParams>(v, r), c[i]);
      head->claim = v;
    }
    if (zc_z) {  // P_{j+1} = P_j (r z + (1 - r)(1 - z))
      const Fr z = *zc_z, one = fp_one<FrParams>();
      const Fr e = fp_add<FrParams>(fp_mul<FrParams>(r, z), fp_mul<FrParams>(fp_sub<FrParams>(one, r), fp_sub<FrParams>(one, z)));
      head->zc_prefix_prev = zc_P;
      head->zc_prefix = fp_mul<FrParams>(zc_P, e);
    }
  }
  if (foldc && t < 8) fold_table_row(s_chal, t, foldc);
}

}  // namespace qz
