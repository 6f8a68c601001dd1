// BN254 Fr / Fq arithmetic for sm_100a: 8x32-bit limbs, Montgomery form (R = 2^256), values always in [0, p).
//
// This is the arithmetic ark-ff 0.5.0 supplies to the reference (Fr at hyperplonk/src/piops/sumcheck.rs:53-92,
// Fq under ark-ec's group law used by pcs/src/kzg.rs:72).  Byte layout of an element equals arkworks' in-memory
// layout (4 x u64 little-endian Montgomery limbs == 8 x u32), so host buffers cross the C ABI without conversion.
//
// Multiplication is a word-serial Montgomery product on the 32-bit integer multiply pipe.  Per multiplier word b_i the
// partial products a_j*b_i with j even occupy disjoint 64-bit lanes (columns j, j+1), as do the ones with j odd, so
// each half-row is one carry chain of mad.lo.cc / madc.hi.cc pairs with no per-product carry fix-up; ptxas fuses
// each lo/hi pair into a single IMAD.WIDE.U32 with predicate carry.  Because p < 2^254 the running value stays
// below 2p < 2^255, so the accumulator never needs a tenth word and the final reduction is one conditional subtract.
#pragma once
#include <cstdint>
#include "ff_consts.cuh"

namespace qz {

#define QZ_DEV __device__ __forceinline__

// programmatic dependent launch (ctx.cuh QZ_LAUNCH_PDL): let the next kernel of the stream be scheduled / wait until
// everything the previous kernel wrote is visible.  Both are no-ops under an ordinary launch.
QZ_DEV void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
QZ_DEV void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class P>
struct Fp {
  uint32_t v[8];
};

// one half-row:  t[0..7] += {a0,a1,a2,a3} (each a 32-bit word occupying lane (2i, 2i+1)) * b ;  carry -> t8
//   last == false:  t8 += carry   (t8 may be non-zero)
QZ_DEV void mad_chain_even(uint32_t& t0, uint32_t& t1, uint32_t& t2, uint32_t& t3, uint32_t& t4, uint32_t& t5,
                           uint32_t& t6, uint32_t& t7, uint32_t& t8, uint32_t a0, uint32_t a1, uint32_t a2,
                           uint32_t a3, uint32_t b) {
  asm volatile(
      "mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;\n\t"
      : "+r"(t0), "+r"(t1), "+r"(t2), "+r"(t3), "+r"(t4), "+r"(t5), "+r"(t6), "+r"(t7), "+r"(t8)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// r = a - p if a >= p else a   (a < 2p)
template <class P>
QZ_DEV void fp_reduce_once(uint32_t r[8], const uint32_t a[8]) {
  uint32_t d[8], borrow;
  asm volatile(
      "sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;\n\t"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(borrow)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(P::MOD(0)),
        "r"(P::MOD(1)), "r"(P::MOD(2)), "r"(P::MOD(3)), "r"(P::MOD(4)), "r"(P::MOD(5)), "r"(P::MOD(6)),
        "r"(P::MOD(7)));
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = borrow ? a[i] : d[i];
}

// ---- shift-free multiplier -------------------------------------------------------------------------------------------
// A word-serial multiplier that shifts its accumulator by one word per row (tools/ff_variants.cuh, fp_mul_inline) costs
// ~136 MOVs per product, because IMAD.WIDE needs even-aligned register pairs.  This one keeps TWO accumulators, E holding the 64-bit lanes
// that start at even columns and O the lanes that start at odd columns (value = E + O * 2^32).  Dividing by 2^32 after a
// row turns O into the new E for free and E's upper three lanes into the new O's lower three; that one-lane move is
// folded into the next row's multiply-add (destination lane j = product + old lane j+1).  E's stray low word (column
// 0 of the new frame) is added to the new E and its carry enters the new O's chain, which starts exactly one column up.
QZ_DEV void mul4(uint32_t* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
  asm volatile(
      "mul.lo.u32 %0, %8, %12;\n\t"
      "mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t"
      "mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t"
      "mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t"
      "mul.hi.u32 %7, %11, %12;\n\t"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// d[0..7] += {a0..a3} * b as four 64-bit lanes in one carry chain; the carry out of the top lane is added to `top`
QZ_DEV void cmad4_top(uint32_t* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b, uint32_t& top) {
  asm volatile(
      "mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;\n\t"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7]), "+r"(top)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// same, for the accumulator whose top lane ends at the highest column: no carry can leave it
QZ_DEV void cmad4(uint32_t* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
  asm volatile(
      "mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.u32 %7, %11, %12, %7;\n\t"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// e0 += x[1] (the stray word), then x lane j = {a0..a3}[j] * b + old x lane j+1 (+ carry), top lane = product + carry
QZ_DEV void madc4_rshift(uint32_t* x, uint32_t& e0, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
  asm volatile(
      "add.cc.u32 %8, %8, %1;\n\t"
      "madc.lo.cc.u32 %0, %9, %13, %2;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %4;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %6;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, 0;\n\t"
      "madc.hi.u32 %7, %12, %13, 0;\n\t"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(e0)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

template <class P>
QZ_DEV void mont_reduce_row(uint32_t* ev, uint32_t* od) {  // ev: even lanes (column 0 first), od: odd lanes
  const uint32_t m = ev[0] * P::INV;
  cmad4(od, P::MOD(1), P::MOD(3), P::MOD(5), P::MOD(7), m);
  cmad4_top(ev, P::MOD(0), P::MOD(2), P::MOD(4), P::MOD(6), m, od[7]);
}

template <class P>
QZ_DEV Fp<P> fp_mul_v1(const Fp<P>& a, const Fp<P>& b) {
  uint32_t X[8], Y[8];
  mul4(X, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
  mul4(Y, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
  mont_reduce_row<P>(X, Y);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    // even lanes X, odd lanes Y  ->  even lanes Y, odd lanes X
    madc4_rshift(X, Y[0], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
    cmad4_top(Y, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i], X[7]);
    mont_reduce_row<P>(Y, X);
    if (i + 1 < 8) {
      madc4_rshift(Y, X[0], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
      cmad4_top(X, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1], Y[7]);
      mont_reduce_row<P>(X, Y);
    }
  }
  // after rows 0..7 the even lanes are in Y (Y[0] == 0) and the odd lanes in X: result = X + Y[1..7]
  uint32_t t[8];
  asm volatile(
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;\n\t"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
      : "r"(X[0]), "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]), "r"(Y[1]), "r"(Y[2]),
        "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
  Fp<P> r;
  fp_reduce_once<P>(r.v, t);
  return r;
}

// a*b + c*d with ONE Montgomery reduction: both product rows enter the running sum before the reduction row, so the
// pair costs 8 x 24 = 192 multiply-adds instead of 256.  With a, c < p the running sum stays below 3p < 2^256
// (t' < (t + 3p(2^32 - 1)) / 2^32), so the same nine columns suffice and two conditional subtractions finish.
template <class P>
QZ_DEV Fp<P> fp_mul2_add(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
  uint32_t X[8], Y[8];
  mul4(X, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
  mul4(Y, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
  cmad4(Y, c.v[1], c.v[3], c.v[5], c.v[7], d.v[0]);
  cmad4_top(X, c.v[0], c.v[2], c.v[4], c.v[6], d.v[0], Y[7]);
  mont_reduce_row<P>(X, Y);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    madc4_rshift(X, Y[0], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
    cmad4_top(Y, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i], X[7]);
    cmad4(X, c.v[1], c.v[3], c.v[5], c.v[7], d.v[i]);
    cmad4_top(Y, c.v[0], c.v[2], c.v[4], c.v[6], d.v[i], X[7]);
    mont_reduce_row<P>(Y, X);
    if (i + 1 < 8) {
      madc4_rshift(Y, X[0], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
      cmad4_top(X, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1], Y[7]);
      cmad4(Y, c.v[1], c.v[3], c.v[5], c.v[7], d.v[i + 1]);
      cmad4_top(X, c.v[0], c.v[2], c.v[4], c.v[6], d.v[i + 1], Y[7]);
      mont_reduce_row<P>(X, Y);
    }
  }
  uint32_t t[8], u[8];
  asm volatile(
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;\n\t"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
      : "r"(X[0]), "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]), "r"(Y[1]), "r"(Y[2]),
        "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
  Fp<P> r;
  fp_reduce_once<P>(u, t);
  fp_reduce_once<P>(r.v, u);
  return r;
}

// The multiplier is ~350 SASS instructions (5.6 KB); a mixed point addition inlines 10 of them.  ncu (r01) shows
// "no_instruction" as the top stall of msm_accumulate, but calling ONE out-of-line copy per field (-DQZ_OUTLINE_MUL)
// measured no faster for the MSM (63.3 vs 62.2 ms at 2^24) and 14% slower for the sumcheck rounds, so inlining stays
// the default.
#ifdef QZ_OUTLINE_MUL
template <class P>
__device__ __noinline__ Fp<P> fp_mul(const Fp<P> a, const Fp<P> b) {
  return fp_mul_v1<P>(a, b);
}
#else
template <class P>
QZ_DEV Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
  return fp_mul_v1<P>(a, b);  // 542 vs 610 SMSP-cycles per warp-product in tools/mulbench.cu (floor: 528)
}
#endif

// ---- squaring -----------------------------------------------------------------------------------------------------------
// a^2 with the interleaved multiplier.  A dedicated square (36 limb products, reduction afterwards: 100 IMAD.WIDE instead
// of 136) was built and measured SLOWER in msm_accumulate on B200 (31.05 vs 30.33 ms at 2^24): its eight reduction-only
// rows form one dependent chain with no product rows to overlap.  The code lives on as tools/ff_variants.cuh.
template <class P>
QZ_DEV Fp<P> fp_sqr(const Fp<P>& a) {
  return fp_mul<P>(a, a);
}

template <class P>
QZ_DEV Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
  uint32_t s[8];
  asm volatile(
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  Fp<P> r;
  fp_reduce_once<P>(r.v, s);
  return r;
}

template <class P>
QZ_DEV Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
  uint32_t d[8], mask;
  asm volatile(
      "sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;\n\t"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(mask)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // mask = 0xffffffff when a < b: add p back
  Fp<P> r;
  asm volatile(
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7])
      : "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
        "r"(P::MOD(0) & mask), "r"(P::MOD(1) & mask), "r"(P::MOD(2) & mask), "r"(P::MOD(3) & mask),
        "r"(P::MOD(4) & mask), "r"(P::MOD(5) & mask), "r"(P::MOD(6) & mask), "r"(P::MOD(7) & mask));
  return r;
}

template <class P>
QZ_DEV Fp<P> fp_zero() {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}
template <class P>
QZ_DEV Fp<P> fp_one() {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = P::ONE(i);
  return r;
}
template <class P>
QZ_DEV bool fp_is_zero(const Fp<P>& a) {
  return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}
template <class P>
QZ_DEV bool fp_eq(const Fp<P>& a, const Fp<P>& b) {
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x |= a.v[i] ^ b.v[i];
  return x == 0;
}
template <class P>
QZ_DEV Fp<P> fp_neg(const Fp<P>& a) {
  return fp_sub<P>(fp_zero<P>(), a);
}
template <class P>
QZ_DEV Fp<P> fp_dbl(const Fp<P>& a) {
  return fp_add<P>(a, a);
}
// ---- deferred reduction -------------------------------------------------------------------------------------------------
// A sum of products  sum_i a_i * b_i  that is only ever needed mod p can skip the Montgomery reduction of every term:
// the 512-bit products (64 IMAD.WIDE each instead of 128) are added into a 17-word accumulator and reduced once.
// 17 words hold 2^36 products of values < p < 2^254.
struct FpWide {
  uint32_t w[17];
};
QZ_DEV void wide_zero(FpWide& acc) {
#pragma unroll
  for (int i = 0; i < 17; i++) acc.w[i] = 0;
}
// acc += a * b  (plain integer product of the two Montgomery residues)
template <class P>
QZ_DEV void wide_mul_acc(FpWide& acc, const Fp<P>& a, const Fp<P>& b) {
  // E[k] holds column k in 64-bit lanes that start at even columns, O[k] column k+1 in lanes that start at odd columns;
  // a row's carry out of its top lane lands in a word no earlier row has filled with more than a carry.
  uint32_t E[17], O[15];
#pragma unroll
  for (int i = 0; i < 17; i++) E[i] = 0;
#pragma unroll
  for (int i = 0; i < 15; i++) O[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    cmad4_top(&E[i], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i], E[i + 8]);
    cmad4_top(&O[i], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i], O[i + 8]);
    cmad4_top(&O[i], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1], O[i + 8]);
    cmad4_top(&E[i + 2], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1], E[i + 10]);
  }
  // acc += E + (O << 32); the product is below 2^512, so E[16] is zero and O ends at column 15
  asm volatile(
      "add.cc.u32 %0, %0, %17;\n\t"
      "addc.cc.u32 %1, %1, %18;\n\t"
      "addc.cc.u32 %2, %2, %19;\n\t"
      "addc.cc.u32 %3, %3, %20;\n\t"
      "addc.cc.u32 %4, %4, %21;\n\t"
      "addc.cc.u32 %5, %5, %22;\n\t"
      "addc.cc.u32 %6, %6, %23;\n\t"
      "addc.cc.u32 %7, %7, %24;\n\t"
      "addc.cc.u32 %8, %8, %25;\n\t"
      "addc.cc.u32 %9, %9, %26;\n\t"
      "addc.cc.u32 %10, %10, %27;\n\t"
      "addc.cc.u32 %11, %11, %28;\n\t"
      "addc.cc.u32 %12, %12, %29;\n\t"
      "addc.cc.u32 %13, %13, %30;\n\t"
      "addc.cc.u32 %14, %14, %31;\n\t"
      "addc.cc.u32 %15, %15, %32;\n\t"
      "addc.u32 %16, %16, 0;\n\t"
      : "+r"(acc.w[0]), "+r"(acc.w[1]), "+r"(acc.w[2]), "+r"(acc.w[3]), "+r"(acc.w[4]), "+r"(acc.w[5]), "+r"(acc.w[6]),
        "+r"(acc.w[7]), "+r"(acc.w[8]), "+r"(acc.w[9]), "+r"(acc.w[10]), "+r"(acc.w[11]), "+r"(acc.w[12]),
        "+r"(acc.w[13]), "+r"(acc.w[14]), "+r"(acc.w[15]), "+r"(acc.w[16])
      : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
        "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]));
  asm volatile(
      "add.cc.u32 %0, %0, %16;\n\t"
      "addc.cc.u32 %1, %1, %17;\n\t"
      "addc.cc.u32 %2, %2, %18;\n\t"
      "addc.cc.u32 %3, %3, %19;\n\t"
      "addc.cc.u32 %4, %4, %20;\n\t"
      "addc.cc.u32 %5, %5, %21;\n\t"
      "addc.cc.u32 %6, %6, %22;\n\t"
      "addc.cc.u32 %7, %7, %23;\n\t"
      "addc.cc.u32 %8, %8, %24;\n\t"
      "addc.cc.u32 %9, %9, %25;\n\t"
      "addc.cc.u32 %10, %10, %26;\n\t"
      "addc.cc.u32 %11, %11, %27;\n\t"
      "addc.cc.u32 %12, %12, %28;\n\t"
      "addc.cc.u32 %13, %13, %29;\n\t"
      "addc.cc.u32 %14, %14, %30;\n\t"
      "addc.u32 %15, %15, 0;\n\t"
      : "+r"(acc.w[1]), "+r"(acc.w[2]), "+r"(acc.w[3]), "+r"(acc.w[4]), "+r"(acc.w[5]), "+r"(acc.w[6]), "+r"(acc.w[7]),
        "+r"(acc.w[8]), "+r"(acc.w[9]), "+r"(acc.w[10]), "+r"(acc.w[11]), "+r"(acc.w[12]), "+r"(acc.w[13]),
        "+r"(acc.w[14]), "+r"(acc.w[15]), "+r"(acc.w[16])
      : "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]),
        "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
}
// (sum of products) * R^-1 mod p, i.e. the sum of the Montgomery products:  T = lo + mid R + top R^2  ->  lo R^-1 + mid + top R
template <class P>
QZ_DEV Fp<P> wide_reduce(const FpWide& acc) {
  Fp<P> lo, mid, top = fp_zero<P>(), raw_one = fp_zero<P>(), r2;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    lo.v[i] = acc.w[i];
    mid.v[i] = acc.w[8 + i];
    r2.v[i] = P::R2(i);
  }
  top.v[0] = acc.w[16];
  raw_one.v[0] = 1;
  // fp_mul's first operand must be < p, the second may be any 256-bit value
  return fp_add<P>(fp_add<P>(fp_mul<P>(raw_one, lo), fp_mul<P>(fp_one<P>(), mid)), fp_mul<P>(r2, top));
}

// ---- multiplication by a FIXED element through precomputed shifted multiples -----------------------------------------------
// A sumcheck pass folds every table entry with the SAME challenge r (lo' = a0 + r (a1 - a0), sumcheck.rs:81-92), 2 * k
// products per pair -- more than half of a round's multiplier work.  For a fixed r the products r * 2^(32 i) mod p can be
// prepared once per round, and then  r * d = sum_i d_i * (r 2^(32 i) mod p)  needs no interleaved reduction: eight rows
// of eight multiply-adds with all rows aligned at column 0, plus THREE reduction rows to bring the 288-bit sum back
// below p -- 88 multiply-adds instead of 128 + 8.  The table holds C_i = r * 2^(32 i + 32 q_i) mod p (canonical, < p)
// with q_i = the number of reduction rows (each a division by 2^32) applied after row i enters the sum: rows 0..3
// enter first (q = 3), one reduction, rows 4..7 (q = 2), two reductions.  Bounds (p < 0.19 * 2^256, every d_i < 2^32
// whatever d is, so d may be any 256-bit value): after rows 0..3 the sum is < 2^34 p < 0.76 * 2^288; + m p < 0.95 * 2^288
// (nine columns suffice, as in fp_mul_v1); / 2^32 -> < 5 p; + rows 4..7 -> < 0.76 * 2^288 + 5 p; reduced twice ->
// < p (1 + 2^-29): one conditional subtraction gives the canonical result.  fold_consts_exponent(i) = 32 i + 32 q_i.
QZ_DEV void rshift_only(uint32_t* x, uint32_t& e0) {  // running sum / 2^32 without a product row (see madc4_rshift)
  asm volatile(
      "add.cc.u32 %8, %8, %1;\n\t"
      "addc.cc.u32 %0, %2, 0;\n\t"
      "addc.cc.u32 %1, %3, 0;\n\t"
      "addc.cc.u32 %2, %4, 0;\n\t"
      "addc.cc.u32 %3, %5, 0;\n\t"
      "addc.cc.u32 %4, %6, 0;\n\t"
      "addc.cc.u32 %5, %7, 0;\n\t"
      "addc.u32 %6, 0, 0;\n\t"
      "mov.u32 %7, 0;\n\t"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(e0));
}
__host__ __device__ constexpr int fold_consts_exponent(int i) { return 32 * i + (i < 4 ? 96 : 64); }
// C: 64 words, C[8 i + j] = word j of C_i (constant memory: the words become instruction operands, no registers held)
template <class P>
QZ_DEV Fp<P> fp_mul_fixed(const Fp<P>& d, const uint32_t* C) {
  uint32_t X[8], Y[8];
  mul4(X, C[0], C[2], C[4], C[6], d.v[0]);
  mul4(Y, C[1], C[3], C[5], C[7], d.v[0]);
#pragma unroll
  for (int i = 1; i < 4; i++) {
    cmad4(Y, C[8 * i + 1], C[8 * i + 3], C[8 * i + 5], C[8 * i + 7], d.v[i]);
    cmad4_top(X, C[8 * i + 0], C[8 * i + 2], C[8 * i + 4], C[8 * i + 6], d.v[i], Y[7]);
  }
  mont_reduce_row<P>(X, Y);  // X[0] == 0; even lanes X, odd lanes Y -> (shift) even lanes Y, odd lanes X
  madc4_rshift(X, Y[0], C[33], C[35], C[37], C[39], d.v[4]);
  cmad4_top(Y, C[32], C[34], C[36], C[38], d.v[4], X[7]);
#pragma unroll
  for (int i = 5; i < 8; i++) {
    cmad4(X, C[8 * i + 1], C[8 * i + 3], C[8 * i + 5], C[8 * i + 7], d.v[i]);
    cmad4_top(Y, C[8 * i + 0], C[8 * i + 2], C[8 * i + 4], C[8 * i + 6], d.v[i], X[7]);
  }
  mont_reduce_row<P>(Y, X);  // Y[0] == 0
  rshift_only(Y, X[0]);      // even lanes X, odd lanes Y
  mont_reduce_row<P>(X, Y);  // X[0] == 0: the value / 2^32 is Y + X[1..7]
  uint32_t t[8];
  asm volatile(
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;\n\t"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
      : "r"(Y[0]), "r"(Y[1]), "r"(Y[2]), "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]), "r"(X[1]), "r"(X[2]),
        "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]));
  Fp<P> r;
  fp_reduce_once<P>(r.v, t);
  return r;
}
// a - b + p without the conditional correction: a value in (0, 2p) congruent to a - b -- good enough for an operand
// whose magnitude does not matter (fp_mul's second operand, fp_mul_fixed's d)
template <class P>
QZ_DEV Fp<P> fp_sub_lazy(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  asm volatile(
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(P::MOD(0)), "r"(P::MOD(1)), "r"(P::MOD(2)), "r"(P::MOD(3)), "r"(P::MOD(4)), "r"(P::MOD(5)), "r"(P::MOD(6)),
        "r"(P::MOD(7)));
  asm volatile(
      "sub.cc.u32 %0, %0, %8;\n\t"
      "subc.cc.u32 %1, %1, %9;\n\t"
      "subc.cc.u32 %2, %2, %10;\n\t"
      "subc.cc.u32 %3, %3, %11;\n\t"
      "subc.cc.u32 %4, %4, %12;\n\t"
      "subc.cc.u32 %5, %5, %13;\n\t"
      "subc.cc.u32 %6, %6, %14;\n\t"
      "subc.u32 %7, %7, %15;\n\t"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
      : "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  return r;
}

// plain 256-bit integer addition / subtraction (no reduction; the caller guarantees 0 <= result < 2^256): for operands
// whose magnitude does not matter -- wide_mul_acc takes any 256-bit values, fp_mul any 256-bit second operand
template <class P>
QZ_DEV Fp<P> u256_add(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  asm volatile(
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  return r;
}
template <class P>
QZ_DEV Fp<P> u256_p_minus(const Fp<P>& a) {  // p - a for a <= p
  Fp<P> r;
  asm volatile(
      "sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32 %7, %15, %23;\n\t"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(P::MOD(0)), "r"(P::MOD(1)), "r"(P::MOD(2)), "r"(P::MOD(3)), "r"(P::MOD(4)), "r"(P::MOD(5)), "r"(P::MOD(6)),
        "r"(P::MOD(7)), "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
        "r"(a.v[7]));
  return r;
}

// Montgomery -> canonical limbs (multiply by 1)
template <class P>
QZ_DEV Fp<P> fp_from_mont(const Fp<P>& a) {
  Fp<P> one_raw = fp_zero<P>();
  one_raw.v[0] = 1;
  return fp_mul<P>(a, one_raw);
}
// canonical limbs (any value < 2^256, not necessarily < p) -> Montgomery, fully reduced.
// fp_mul needs its FIRST operand < p (it bounds the running sum); the second may be any 256-bit value.
template <class P>
QZ_DEV Fp<P> fp_to_mont(const Fp<P>& a) {
  Fp<P> r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.v[i] = P::R2(i);
  return fp_mul<P>(r2, a);
}
// a^(p-2); 0 -> 0.  Rare path (affine conversion, interpolation constants): plain square-and-multiply.
template <class P>
__device__ __noinline__ Fp<P> fp_inv(const Fp<P>& a) {
  Fp<P> acc = fp_one<P>();
  for (int i = 255; i >= 0; i--) {
    acc = fp_sqr<P>(acc);
    if ((P::PM2(i >> 5) >> (i & 31)) & 1) acc = fp_mul<P>(acc, a);
  }
  return acc;
}

// The same inverse by the binary extended Euclidean algorithm, for the places where ONE thread inverts ONE element on
// the critical path (affine conversion at the end of an MSM, the MLPCS challenge, the zero-check claim): about 500
// shift steps and 350 subtract steps of 256-bit integer work instead of 380 dependent multiplications -- measured
// ~4x lower latency.  Data-dependent control flow, so bulk callers (one inversion per thread across a warp) keep
// fp_inv.  Invariants: x1 * a = u and x2 * a = v (mod p); u, v odd positive with gcd 1 on entry to every subtract.
// The input is a Montgomery residue aR, the raw inverse is a^-1 R^-1, and one product with R^3 restores the form.
template <class P>
__device__ __noinline__ Fp<P> fp_inv_serial(const Fp<P>& a) {
  if (fp_is_zero<P>(a)) return a;
  uint32_t u[8], v[8];
  Fp<P> x1 = fp_zero<P>(), x2 = fp_zero<P>();
  x1.v[0] = 1;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u[i] = a.v[i];
    v[i] = P::MOD(i);
  }
  auto is_one = [](const uint32_t* x) { return x[0] == 1 && (x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7]) == 0; };
  auto shr1 = [](uint32_t* x) {
#pragma unroll
    for (int i = 0; i < 7; i++) x[i] = __funnelshift_r(x[i], x[i + 1], 1);
    x[7] >>= 1;
  };
  auto halve = [&](Fp<P>& x) {  // x / 2 mod p: (x + p) / 2 when x is odd; x + p < 2^255 never carries out
    if (x.v[0] & 1) {
      asm volatile(
          "add.cc.u32 %0, %0, %8;\n\t"
          "addc.cc.u32 %1, %1, %9;\n\t"
          "addc.cc.u32 %2, %2, %10;\n\t"
          "addc.cc.u32 %3, %3, %11;\n\t"
          "addc.cc.u32 %4, %4, %12;\n\t"
          "addc.cc.u32 %5, %5, %13;\n\t"
          "addc.cc.u32 %6, %6, %14;\n\t"
          "addc.u32 %7, %7, %15;\n\t"
          : "+r"(x.v[0]), "+r"(x.v[1]), "+r"(x.v[2]), "+r"(x.v[3]), "+r"(x.v[4]), "+r"(x.v[5]), "+r"(x.v[6]), "+r"(x.v[7])
          : "r"(P::MOD(0)), "r"(P::MOD(1)), "r"(P::MOD(2)), "r"(P::MOD(3)), "r"(P::MOD(4)), "r"(P::MOD(5)), "r"(P::MOD(6)),
            "r"(P::MOD(7)));
    }
    shr1(x.v);
  };
  // d = x - y, returns the borrow (0xffffffff when x < y)
  auto sub_borrow = [](uint32_t* d, const uint32_t* x, const uint32_t* y) {
    uint32_t borrow;
    asm volatile(
        "sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(borrow)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(y[0]), "r"(y[1]),
          "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
    return borrow;
  };
#pragma unroll 1
  while (!is_one(u) && !is_one(v)) {
#pragma unroll 1
    while (!(u[0] & 1)) {
      shr1(u);
      halve(x1);
    }
#pragma unroll 1
    while (!(v[0] & 1)) {
      shr1(v);
      halve(x2);
    }
    uint32_t d[8];
    if (sub_borrow(d, u, v) == 0) {  // u >= v
#pragma unroll
      for (int i = 0; i < 8; i++) u[i] = d[i];
      x1 = fp_sub<P>(x1, x2);
    } else {
      sub_borrow(v, v, u);
      x2 = fp_sub<P>(x2, x1);
    }
  }
  Fp<P> r3;
#pragma unroll
  for (int i = 0; i < 8; i++) r3.v[i] = P::R3(i);
  return fp_mul<P>(is_one(u) ? x1 : x2, r3);
}

// 128-bit vectorised global access: an element is two uint4
template <class P>
QZ_DEV Fp<P> fp_load(const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = q[0], hi = q[1];
  Fp<P> r;
  r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
  r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
  return r;
}
template <class P>
QZ_DEV void fp_store(void* p, const Fp<P>& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace qz
