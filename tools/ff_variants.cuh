// Recorded NEGATIVE results and bench-only variants of the field arithmetic, kept out of the product headers
// (quill_zkvm_b200/csrc/ff.cuh) so that tools/mulbench.cu can still measure them:
//   fp_mul_inline   the first word-serial Montgomery multiplier: one accumulator shifted by a word per row -- 610 SMSP-cycles per
//                   warp-product against 542 for the shift-free fp_mul_v1 (136 MOVs per product to re-align IMAD.WIDE pairs)
//   fp_sqr_sos      dedicated squaring, 100 IMAD.WIDE instead of 136 -- measured slower inside msm_accumulate (31.05 vs 30.33 ms)
#pragma once
#include "../quill_zkvm_b200/csrc/ff.cuh"

namespace qz {
// the odd half-row ends exactly at the top word: no carry can leave it (value < 2^288)
QZ_DEV void mad_chain_odd(uint32_t& t1, uint32_t& t2, uint32_t& t3, uint32_t& t4, uint32_t& t5, uint32_t& t6,
                          uint32_t& t7, uint32_t& t8, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                          uint32_t b) {
  asm volatile(
      "mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.u32 %7, %11, %12, %7;\n\t"
      : "+r"(t1), "+r"(t2), "+r"(t3), "+r"(t4), "+r"(t5), "+r"(t6), "+r"(t7), "+r"(t8)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}


template <class P>
QZ_DEV Fp<P> fp_mul_inline(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0, t8 = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t bi = b.v[i];
    mad_chain_even(t0, t1, t2, t3, t4, t5, t6, t7, t8, a.v[0], a.v[2], a.v[4], a.v[6], bi);
    mad_chain_odd(t1, t2, t3, t4, t5, t6, t7, t8, a.v[1], a.v[3], a.v[5], a.v[7], bi);
    const uint32_t m = t0 * P::INV;
    mad_chain_even(t0, t1, t2, t3, t4, t5, t6, t7, t8, P::MOD(0), P::MOD(2), P::MOD(4), P::MOD(6), m);
    mad_chain_odd(t1, t2, t3, t4, t5, t6, t7, t8, P::MOD(1), P::MOD(3), P::MOD(5), P::MOD(7), m);
    // t0 == 0 now: shift down one word
    t0 = t1; t1 = t2; t2 = t3; t3 = t4; t4 = t5; t5 = t6; t6 = t7; t7 = t8; t8 = 0;
  }
  const uint32_t t[8] = {t0, t1, t2, t3, t4, t5, t6, t7};
  Fp<P> r;
  fp_reduce_once<P>(r.v, t);
  return r;
}



QZ_DEV void cmad1_top(uint32_t* d, uint32_t a0, uint32_t b, uint32_t& top) {
  asm volatile(
      "mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      : "+r"(d[0]), "+r"(d[1]), "+r"(top)
      : "r"(a0), "r"(b));
}
QZ_DEV void cmad2_top(uint32_t* d, uint32_t a0, uint32_t a1, uint32_t b, uint32_t& top) {
  asm volatile(
      "mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32 %4, %4, 0;\n\t"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(top)
      : "r"(a0), "r"(a1), "r"(b));
}
QZ_DEV void cmad3_top(uint32_t* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b, uint32_t& top) {
  asm volatile(
      "mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
      "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
      "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
      "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
      "addc.u32 %6, %6, 0;\n\t"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(top)
      : "r"(a0), "r"(a1), "r"(a2), "r"(b));
}
// division of the running sum by 2^32 without a product row: x (the lanes whose low column was just cancelled)
// becomes the odd-lane accumulator of the new frame, its stray word joins the new even accumulator's word 0
template <class P>
QZ_DEV Fp<P> fp_sqr_sos(const Fp<P>& a) {
  // 1. off-diagonal products: E[k] = column k (lanes at even columns), O[k] = column k + 1 (lanes at odd columns)
  uint32_t E[16], O[16];
#pragma unroll
  for (int i = 0; i < 16; i++) E[i] = O[i] = 0;
  cmad4_top(&O[0], a.v[1], a.v[3], a.v[5], a.v[7], a.v[0], O[8]);   // row 0: columns 1, 3, 5, 7
  cmad3_top(&E[2], a.v[2], a.v[4], a.v[6], a.v[0], E[8]);           //        columns 2, 4, 6
  cmad3_top(&O[2], a.v[2], a.v[4], a.v[6], a.v[1], O[8]);           // row 1: columns 3, 5, 7
  cmad3_top(&E[4], a.v[3], a.v[5], a.v[7], a.v[1], E[10]);          //        columns 4, 6, 8
  cmad3_top(&O[4], a.v[3], a.v[5], a.v[7], a.v[2], O[10]);          // row 2: columns 5, 7, 9
  cmad2_top(&E[6], a.v[4], a.v[6], a.v[2], E[10]);                  //        columns 6, 8
  cmad2_top(&O[6], a.v[4], a.v[6], a.v[3], O[10]);                  // row 3: columns 7, 9
  cmad2_top(&E[8], a.v[5], a.v[7], a.v[3], E[12]);                  //        columns 8, 10
  cmad2_top(&O[8], a.v[5], a.v[7], a.v[4], O[12]);                  // row 4: columns 9, 11
  cmad1_top(&E[10], a.v[6], a.v[4], E[12]);                         //        column 10
  cmad1_top(&O[10], a.v[6], a.v[5], O[12]);                         // row 5: column 11
  cmad1_top(&E[12], a.v[7], a.v[5], E[14]);                         //        column 12
  cmad1_top(&O[12], a.v[7], a.v[6], O[14]);                         // row 6: column 13
  // S = E + (O << 32), columns 1..15 (column 0 is empty)
  uint32_t S[16];
  S[0] = 0;
  asm volatile(
      "add.cc.u32 %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32 %14, %29, %44;\n\t"
      : "=r"(S[1]), "=r"(S[2]), "=r"(S[3]), "=r"(S[4]), "=r"(S[5]), "=r"(S[6]), "=r"(S[7]), "=r"(S[8]), "=r"(S[9]),
        "=r"(S[10]), "=r"(S[11]), "=r"(S[12]), "=r"(S[13]), "=r"(S[14]), "=r"(S[15])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]), "r"(E[10]),
        "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]),
        "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
  // T = 2S + sum_i a_i^2 2^(64 i)
  uint32_t T[16];
  T[0] = 0;
#pragma unroll
  for (int k = 1; k < 16; k++) T[k] = __funnelshift_l(S[k - 1], S[k], 1);
  asm volatile(
      "mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
      "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
      "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
      "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
      "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
      "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
      "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
      "madc.hi.u32 %15, %23, %23, %15;\n\t"
      : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]),
        "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
  // 2. cancel T_lo: eight reduction-only rows (even lanes X, odd lanes Y, roles swapping as in fp_mul_v1)
  uint32_t X[8], Y[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    X[i] = T[i];
    Y[i] = 0;
  }
  mont_reduce_row<P>(X, Y);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    rshift_only(X, Y[0]);
    mont_reduce_row<P>(Y, X);
    if (i + 1 < 8) {
      rshift_only(Y, X[0]);
      mont_reduce_row<P>(X, Y);
    }
  }
  // U = X + Y[1..7]  (<= p),  result = U + T_hi  (< 2p)
  uint32_t u[8], t[8];
  asm volatile(
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;\n\t"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
      : "r"(X[0]), "r"(X[1]), "r"(X[2]), "r"(X[3]), "r"(X[4]), "r"(X[5]), "r"(X[6]), "r"(X[7]), "r"(Y[1]), "r"(Y[2]),
        "r"(Y[3]), "r"(Y[4]), "r"(Y[5]), "r"(Y[6]), "r"(Y[7]));
  asm volatile(
      "add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
      : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(T[8]), "r"(T[9]),
        "r"(T[10]), "r"(T[11]), "r"(T[12]), "r"(T[13]), "r"(T[14]), "r"(T[15]));
  Fp<P> r;
  fp_reduce_once<P>(r.v, t);
  return r;
}

}  // namespace qz
