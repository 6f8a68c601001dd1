// CPU ORACLE (test infrastructure, NOT product code).  PARITY UNPINNED -- see oracle/README.md.
// BN254 G2 and the optimal-ate pairing, restating what ark-ec / ark-bn254 0.5.0 (Cargo.lock:24-25,36-37) give
// the reference's VERIFIER: `E::pairing` at pcs/src/kzg.rs:98-108 and the G2 half of `KZG::trusted_setup`
// (kzg.rs:49-52).  The pairing is only ever used through an equality of two pairings (a boolean), so any
// non-degenerate bilinear map on the r-torsion gives the verifier's answer; this one is the textbook optimal
// ate on the D-type sextic twist, written for clarity not speed:
//   Fq2 = Fq[i]/(i^2+1),  Fq12 = Fq[w]/(w^12 - 18 w^6 + 82)  (so w^6 = 9 + i = xi),
//   twist psi(x, y) = (x w^2, y w^3) from E'(Fq2): y^2 = x^3 + 3/xi onto E(Fq12): y^2 = x^3 + 3.
// The running point stays in affine Fq2 coordinates; a line through psi(R1), psi(R2) evaluated at P = (xP, yP)
// in G1 is  -yP + (lambda xP) w + (y1 - lambda x1) w^3  with lambda the Fq2 slope on the twist.
#pragma once
#include "g1.hpp"

namespace orc {

struct Fq2 {
  Fq c0, c1;
  static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fq2& o) const { return !(*this == o); }
  Fq2 operator+(const Fq2& o) const { return Fq2{c0 + o.c0, c1 + o.c1}; }
  Fq2 operator-(const Fq2& o) const { return Fq2{c0 - o.c0, c1 - o.c1}; }
  Fq2 neg() const { return Fq2{c0.neg(), c1.neg()}; }
  Fq2 conj() const { return Fq2{c0, c1.neg()}; }
  Fq2 operator*(const Fq2& o) const { return Fq2{c0 * o.c0 - c1 * o.c1, c0 * o.c1 + c1 * o.c0}; }
  Fq2 scale(const Fq& s) const { return Fq2{c0 * s, c1 * s}; }
  Fq2 sqr() const { return *this * *this; }
  Fq2 dbl() const { return *this + *this; }
  Fq2 inverse() const {  // conj / norm; 0 -> 0
    Fq n = (c0.sqr() + c1.sqr()).inverse();
    return Fq2{c0 * n, c1.neg() * n};
  }
  Fq2 pow(const u64 e[4]) const {
    Fq2 acc = one();
    for (int i = 255; i >= 0; i--) {
      acc = acc.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) acc = acc * *this;
    }
    return acc;
  }
};

inline Fq2 xi() { return Fq2{Fq::from_u64(9), Fq::one()}; }
inline Fq2 twist_b() { return xi().inverse().scale(Fq::from_u64(3)); }  // 3 / (9 + i)

struct G2Affine {
  Fq2 x, y;
  bool inf;
  static G2Affine infinity() { return G2Affine{Fq2::zero(), Fq2::zero(), true}; }
  bool on_curve() const { return inf || y.sqr() == x.sqr() * x + twist_b(); }
  G2Affine neg() const { return G2Affine{x, y.neg(), inf}; }
  // affine chord/tangent with the slope handed back (the Miller loop needs it)
  G2Affine add(const G2Affine& o, Fq2* slope = nullptr) const {
    if (inf) return o;
    if (o.inf) return *this;
    Fq2 lam;
    if (x == o.x) {
      if (y != o.y || y.is_zero()) return infinity();
      lam = (x.sqr().dbl() + x.sqr()) * y.dbl().inverse();
    } else {
      lam = (o.y - y) * (o.x - x).inverse();
    }
    if (slope) *slope = lam;
    Fq2 x3 = lam.sqr() - x - o.x;
    return G2Affine{x3, lam * (x - x3) - y, false};
  }
  G2Affine mul_canonical(const u64 k[4]) const {
    G2Affine acc = infinity();
    for (int i = 255; i >= 0; i--) {
      acc = acc.add(acc);
      if ((k[i / 64] >> (i % 64)) & 1) acc = acc.add(*this);
    }
    return acc;
  }
  G2Affine mul(const Fr& s) const {
    u64 k[4];
    s.to_canonical(k);
    return mul_canonical(k);
  }
};

// The standard BN254 G2 generator (EIP-197 / ark-bn254 `G2Affine::generator`), decimal -> limbs at first use.
inline Fq fq_from_decimal(const char* s) {
  Fq acc = Fq::zero(), ten = Fq::from_u64(10);
  for (; *s; s++) acc = acc * ten + Fq::from_u64((u64)(*s - '0'));
  return acc;
}
inline G2Affine g2_generator() {
  return G2Affine{
      Fq2{fq_from_decimal("10857046999023057135944570762232829481370756359578518086990519993285655852781"),
          fq_from_decimal("11559732032986387107991004021392285783925812861821192530917403151452391805634")},
      Fq2{fq_from_decimal("8495653923123431417604973247489272438418190587263600148770280649306958101930"),
          fq_from_decimal("4082367875863433681332203403145435568316851327593401208105741076214120093531")},
      false};
}

struct Fq12 {
  Fq c[12];  // coefficients of w^0 .. w^11
  static Fq12 one() {
    Fq12 r;
    for (auto& v : r.c) v = Fq::zero();
    r.c[0] = Fq::one();
    return r;
  }
  bool operator==(const Fq12& o) const {
    for (int i = 0; i < 12; i++)
      if (c[i] != o.c[i]) return false;
    return true;
  }
  Fq12 operator*(const Fq12& o) const {
    Fq t[23];
    for (auto& v : t) v = Fq::zero();
    for (int i = 0; i < 12; i++) {
      if (c[i].is_zero()) continue;  // the line functions are sparse
      for (int j = 0; j < 12; j++) t[i + j] += c[i] * o.c[j];
    }
    const Fq k18 = Fq::from_u64(18), k82 = Fq::from_u64(82);
    for (int d = 22; d >= 12; d--) {  // w^12 = 18 w^6 - 82
      t[d - 6] += t[d] * k18;
      t[d - 12] -= t[d] * k82;
    }
    Fq12 r;
    for (int i = 0; i < 12; i++) r.c[i] = t[i];
    return r;
  }
  Fq12 pow_bytes_le(const uint8_t* e, size_t n) const {
    Fq12 acc = one();
    for (size_t i = n * 8; i-- > 0;) {
      acc = acc * acc;
      if ((e[i / 8] >> (i % 8)) & 1) acc = acc * *this;
    }
    return acc;
  }
};

// a + b i  placed at w^k:  (a - 9 b) w^k + b w^(k+6)   (i = w^6 - 9)
inline void put_fq2(Fq12& f, int k, const Fq2& v) {
  f.c[k] = v.c0 - v.c1 * Fq::from_u64(9);
  f.c[k + 6] = v.c1;
}

inline Fq12 line_eval(const G2Affine& r1, const Fq2& lam, const G1Affine& p) {
  Fq12 l;
  for (auto& v : l.c) v = Fq::zero();
  l.c[0] = p.y.neg();
  put_fq2(l, 1, lam.scale(p.x));
  put_fq2(l, 3, r1.y - lam * r1.x);
  return l;
}

inline void div_small(const u64 a[4], u64 d, u64 out[4]) {
  u128 rem = 0;
  for (int i = 3; i >= 0; i--) {
    u128 cur = (rem << 64) | a[i];
    out[i] = (u64)(cur / d);
    rem = cur % d;
  }
}

// p-power Frobenius carried to the twist: (x, y) -> (conj(x) xi^((p-1)/3), conj(y) xi^((p-1)/2))
inline G2Affine g2_frobenius(const G2Affine& q) {
  u64 pm1[4], e3[4], e2[4];
  memcpy(pm1, FqTag::MOD, 32);
  pm1[0] -= 1;
  div_small(pm1, 3, e3);
  div_small(pm1, 2, e2);
  return G2Affine{q.x.conj() * xi().pow(e3), q.y.conj() * xi().pow(e2), q.inf};
}

// Miller loop of the optimal ate pairing, loop count 6u+2 = 0x1_9d797039be763ba8 (u = 4965661367192848881)
inline Fq12 miller_loop(const G2Affine& q, const G1Affine& p) {
  if (q.inf || p.inf) return Fq12::one();
  const u64 low = 0x9d797039be763ba8ULL;  // bit 64 is the leading one
  G2Affine r = q;
  Fq12 f = Fq12::one();
  Fq2 lam;
  for (int i = 63; i >= 0; i--) {
    G2Affine r2 = r.add(r, &lam);
    f = f * f * line_eval(r, lam, p);
    r = r2;
    if ((low >> i) & 1) {
      G2Affine rq = r.add(q, &lam);
      f = f * line_eval(r, lam, p);
      r = rq;
    }
  }
  G2Affine q1 = g2_frobenius(q);
  G2Affine nq2 = g2_frobenius(q1).neg();
  G2Affine rq1 = r.add(q1, &lam);
  f = f * line_eval(r, lam, p);
  r = rq1;
  r.add(nq2, &lam);
  f = f * line_eval(r, lam, p);
  return f;
}

// prod_i e(P_i, Q_i) with one final exponentiation; `final_exp` = (q^12 - 1) / r as little-endian bytes
inline Fq12 pairing_product(const G1Affine* ps, const G2Affine* qs, size_t n, const uint8_t* final_exp, size_t exp_len) {
  Fq12 f = Fq12::one();
  for (size_t i = 0; i < n; i++) f = f * miller_loop(qs[i], ps[i]);
  return f.pow_bytes_le(final_exp, exp_len);
}

}  // namespace orc
