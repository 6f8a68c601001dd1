// BN254 G1 (y^2 = x^3 + 3 over Fq) group law for the MSM kernels, in extended Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; ZZ = 0 is the identity).  This replaces the Jacobian arithmetic ark-ec 0.5.0
// performs under VariableBaseMSM::msm_unchecked (called at pcs/src/kzg.rs:72).  The result of an MSM is a group
// element, so the choice of coordinates cannot change the affine point that is finally serialised.
//
// Mixed addition XYZZ + affine costs 8M + 2S (Y3 = R (Q - X3) - Y1 PPP as one fused two-product reduction, i.e. 9.5
// multiplications' worth of multiply-adds) and is the unit of work of bucket accumulation.  Every routine here is
// COMPLETE: the P == +-Q and identity cases are detected (values are canonical, so equality is a limb compare) and
// routed to doubling / identity, because bucket reduction and tiny inputs hit them deterministically.
#pragma once
#include "ff.cuh"

namespace qz {

struct Affine {
  Fq x, y;  // (0, 0) encodes the point at infinity (not on the curve)
};
struct Xyzz {
  Fq x, y, zz, zzz;
};

QZ_DEV bool affine_is_inf(const Affine& a) { return fp_is_zero<FqParams>(a.x) && fp_is_zero<FqParams>(a.y); }
QZ_DEV Affine affine_load(const void* p) {
  Affine a;
  a.x = fp_load<FqParams>(p);
  a.y = fp_load<FqParams>((const uint8_t*)p + 32);
  return a;
}
QZ_DEV void affine_store(void* p, const Affine& a) {
  fp_store<FqParams>(p, a.x);
  fp_store<FqParams>((uint8_t*)p + 32, a.y);
}
QZ_DEV Affine affine_neg(const Affine& a) {
  Affine r;
  r.x = a.x;
  r.y = fp_neg<FqParams>(a.y);  // -0 = 0 keeps the infinity encoding
  return r;
}
QZ_DEV Xyzz xyzz_identity() {
  Xyzz r;
  r.x = fp_zero<FqParams>();
  r.y = fp_zero<FqParams>();
  r.zz = fp_zero<FqParams>();
  r.zzz = fp_zero<FqParams>();
  return r;
}
QZ_DEV bool xyzz_is_identity(const Xyzz& p) { return fp_is_zero<FqParams>(p.zz); }
QZ_DEV Xyzz xyzz_from_affine(const Affine& a) {
  if (affine_is_inf(a)) return xyzz_identity();
  Xyzz r;
  r.x = a.x;
  r.y = a.y;
  r.zz = fp_one<FqParams>();
  r.zzz = fp_one<FqParams>();
  return r;
}
QZ_DEV Xyzz xyzz_neg(const Xyzz& p) {
  Xyzz r = p;
  r.y = fp_neg<FqParams>(p.y);
  return r;
}
QZ_DEV void xyzz_load(Xyzz& p, const void* src) {
  const uint8_t* s = (const uint8_t*)src;
  p.x = fp_load<FqParams>(s);
  p.y = fp_load<FqParams>(s + 32);
  p.zz = fp_load<FqParams>(s + 64);
  p.zzz = fp_load<FqParams>(s + 96);
}
QZ_DEV void xyzz_store(void* dst, const Xyzz& p) {
  uint8_t* d = (uint8_t*)dst;
  fp_store<FqParams>(d, p.x);
  fp_store<FqParams>(d + 32, p.y);
  fp_store<FqParams>(d + 64, p.zz);
  fp_store<FqParams>(d + 96, p.zzz);
}

// 2 * (x, y) for an affine point (mdbl-2008-s-1, a = 0)
QZ_DEV Xyzz xyzz_dbl_affine(const Affine& a) {
  if (affine_is_inf(a)) return xyzz_identity();
  Fq u = fp_dbl<FqParams>(a.y), v = fp_sqr<FqParams>(u), w = fp_mul<FqParams>(u, v), s = fp_mul<FqParams>(a.x, v);
  Fq xx = fp_sqr<FqParams>(a.x), m = fp_add<FqParams>(fp_dbl<FqParams>(xx), xx);
  Xyzz r;
  r.x = fp_sub<FqParams>(fp_sqr<FqParams>(m), fp_dbl<FqParams>(s));
  r.y = fp_mul2_add<FqParams>(m, fp_sub<FqParams>(s, r.x), fp_neg<FqParams>(w), a.y);  // m (s - x3) - w y
  r.zz = v;
  r.zzz = w;
  return r;
}
// dbl-2008-s-1 (a = 0)
QZ_DEV Xyzz xyzz_dbl(const Xyzz& p) {
  if (xyzz_is_identity(p)) return p;
  Fq u = fp_dbl<FqParams>(p.y), v = fp_sqr<FqParams>(u), w = fp_mul<FqParams>(u, v), s = fp_mul<FqParams>(p.x, v);
  Fq xx = fp_sqr<FqParams>(p.x), m = fp_add<FqParams>(fp_dbl<FqParams>(xx), xx);
  Xyzz r;
  r.x = fp_sub<FqParams>(fp_sqr<FqParams>(m), fp_dbl<FqParams>(s));
  r.y = fp_mul2_add<FqParams>(m, fp_sub<FqParams>(s, r.x), fp_neg<FqParams>(w), p.y);  // m (s - x3) - w y
  r.zz = fp_mul<FqParams>(v, p.zz);
  r.zzz = fp_mul<FqParams>(w, p.zzz);
  return r;
}
// madd-2008-s: XYZZ + affine, 8M + 2S, complete
QZ_DEV Xyzz xyzz_add_affine(const Xyzz& p, const Affine& a) {
  if (affine_is_inf(a)) return p;
  if (xyzz_is_identity(p)) return xyzz_from_affine(a);
  Fq u2 = fp_mul<FqParams>(a.x, p.zz), s2 = fp_mul<FqParams>(a.y, p.zzz);
  Fq pp_ = fp_sub<FqParams>(u2, p.x), rr = fp_sub<FqParams>(s2, p.y);
  if (fp_is_zero<FqParams>(pp_)) {
    if (fp_is_zero<FqParams>(rr)) return xyzz_dbl_affine(a);
    return xyzz_identity();
  }
  Fq pp = fp_sqr<FqParams>(pp_), ppp = fp_mul<FqParams>(pp_, pp), q = fp_mul<FqParams>(p.x, pp);
  Xyzz r;
  r.x = fp_sub<FqParams>(fp_sub<FqParams>(fp_sqr<FqParams>(rr), ppp), fp_dbl<FqParams>(q));
  r.y = fp_mul2_add<FqParams>(rr, fp_sub<FqParams>(q, r.x), fp_neg<FqParams>(p.y), ppp);  // r (q - x3) - y1 ppp, one reduction
  r.zz = fp_mul<FqParams>(p.zz, pp);
  r.zzz = fp_mul<FqParams>(p.zzz, ppp);
  return r;
}
// add-2008-s: XYZZ + XYZZ, 12M + 2S, complete
QZ_DEV Xyzz xyzz_add(const Xyzz& p1, const Xyzz& p2) {
  if (xyzz_is_identity(p1)) return p2;
  if (xyzz_is_identity(p2)) return p1;
  Fq u1 = fp_mul<FqParams>(p1.x, p2.zz), u2 = fp_mul<FqParams>(p2.x, p1.zz);
  Fq s1 = fp_mul<FqParams>(p1.y, p2.zzz), s2 = fp_mul<FqParams>(p2.y, p1.zzz);
  Fq pp_ = fp_sub<FqParams>(u2, u1), rr = fp_sub<FqParams>(s2, s1);
  if (fp_is_zero<FqParams>(pp_)) {
    if (fp_is_zero<FqParams>(rr)) return xyzz_dbl(p1);
    return xyzz_identity();
  }
  Fq pp = fp_sqr<FqParams>(pp_), ppp = fp_mul<FqParams>(pp_, pp), q = fp_mul<FqParams>(u1, pp);
  Xyzz r;
  r.x = fp_sub<FqParams>(fp_sub<FqParams>(fp_sqr<FqParams>(rr), ppp), fp_dbl<FqParams>(q));
  r.y = fp_mul2_add<FqParams>(rr, fp_sub<FqParams>(q, r.x), fp_neg<FqParams>(s1), ppp);
  r.zz = fp_mul<FqParams>(fp_mul<FqParams>(p1.zz, p2.zz), pp);
  r.zzz = fp_mul<FqParams>(fp_mul<FqParams>(p1.zzz, p2.zzz), ppp);
  return r;
}
// one inversion: 1/(ZZ*ZZZ) gives both 1/ZZ and 1/ZZZ
static __device__ __noinline__ Affine xyzz_to_affine(const Xyzz& p) {
  Affine a;
  if (xyzz_is_identity(p)) {
    a.x = fp_zero<FqParams>();
    a.y = fp_zero<FqParams>();
    return a;
  }
  Fq inv = fp_inv<FqParams>(fp_mul<FqParams>(p.zz, p.zzz));
  a.x = fp_mul<FqParams>(p.x, fp_mul<FqParams>(inv, p.zzz));
  a.y = fp_mul<FqParams>(p.y, fp_mul<FqParams>(inv, p.zz));
  return a;
}
// the same for the single thread that finishes an MSM: latency-optimised inversion (ff.cuh fp_inv_serial)
static __device__ __noinline__ Affine xyzz_to_affine_serial(const Xyzz& p) {
  Affine a;
  if (xyzz_is_identity(p)) {
    a.x = fp_zero<FqParams>();
    a.y = fp_zero<FqParams>();
    return a;
  }
  Fq inv = fp_inv_serial<FqParams>(fp_mul<FqParams>(p.zz, p.zzz));
  a.x = fp_mul<FqParams>(p.x, fp_mul<FqParams>(inv, p.zzz));
  a.y = fp_mul<FqParams>(p.y, fp_mul<FqParams>(inv, p.zz));
  return a;
}

}  // namespace qz
