// Device-side building blocks of the sumcheck prover: transcript, expression interpreter, block reductions.
// Follows hyperplonk/src/piops/sumcheck.rs:28-114 and transcript/src/transcript.rs:14-75 of the reference.
#pragma once
#include "../../include/quill_b200.h"
#include "blake3.cuh"
#include "ff.cuh"

namespace qz {

constexpr int SC_MAX_DEG = QZ_MAX_ROUND_COEFFS - 1;
constexpr int SC_MAX_COEFFS = QZ_MAX_ROUND_COEFFS;
constexpr int SC_MAX_K = 16;      // tables referenced by one expression
constexpr int SC_MAX_OPS = 512;   // postfix program length
constexpr int SC_MAX_STACK = 16;  // interpreter stack depth
constexpr int SC_MAX_VARS = 40;

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
  uint8_t tstate[32];  // Transcript.state
  Fr r;                // challenge of the last closed round (pending fold)
  Fr evaluation;       // EvaluationClaim.evaluation
};

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
QZ_DEV void sc_round_close(ScHead* head, const Fr* vinv, int d, const Fr* s_evals, Fr* s_coef, uint8_t* s_msg,
                           Fr* out_coeffs_row, uint32_t* out_len, Fr* out_point_slot, int max_coeffs) {
  const int t = threadIdx.x;
  if (t <= d) {
    Fr acc = fp_zero<FrParams>();
    for (int j = 0; j <= d; j++) acc = fp_add<FrParams>(acc, fp_mul<FrParams>(vinv[t * (d + 1) + j], s_evals[j]));
    s_coef[t] = acc;
    out_coeffs_row[t] = acc;
    fr_to_le_bytes(acc, s_msg + 8 + 32 * t);
  } else if (t < max_coeffs) {
    out_coeffs_row[t] = fp_zero<FrParams>();
  }
  __syncthreads();
  if (t == 0) {
    int len = d + 1;
    while (len > 0 && fp_is_zero<FrParams>(s_coef[len - 1])) len--;  // DensePolynomial trims trailing zeros
    *out_len = (uint32_t)len;
    for (int i = 0; i < 8; i++) s_msg[i] = i == 0 ? (uint8_t)len : 0;  // u64 LE length prefix
    tr_absorb(head->tstate, s_msg, 8 + 32 * len);                     // :73
    Fr r = tr_draw_fr(head->tstate);                                   // :77
    head->r = r;
    *out_point_slot = r;
  }
  __syncthreads();
}

}  // namespace qz
