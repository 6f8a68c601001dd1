// CPU ORACLE (test infrastructure, NOT product code).  PARITY UNPINNED -- see oracle/README.md.
// BN254 Fr / Fq in 4x64-bit Montgomery form (R = 2^256), restating what ark-ff 0.5.0 (Cargo.lock:57-58)
// provides to the reference: the in-memory layout of `Fr` (4 little-endian u64 Montgomery limbs), canonical
// little-endian serialization, and `from_le_bytes_mod_order` (transcript/src/transcript.rs:71-75).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {
typedef uint64_t u64;
typedef unsigned __int128 u128;

struct FrTag {
  static constexpr u64 MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL,
                                 0x30644e72e131a029ULL};
};
struct FqTag {
  static constexpr u64 MOD[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL,
                                 0x30644e72e131a029ULL};
};

template <class P>
struct Fp {
  u64 l[4];  // Montgomery limbs, little-endian; value*R mod p, always < p

  struct Consts {
    u64 inv;      // -p^{-1} mod 2^64
    u64 one[4];   // R mod p
    u64 r2[4];    // R^2 mod p
    Consts() {
      u64 x = 1;  // Newton: x = p^{-1} mod 2^64
      for (int i = 0; i < 7; i++) x *= 2 - P::MOD[0] * x;
      inv = (u64)0 - x;
      // R mod p by 256 modular doublings of 1, R^2 by 256 more
      u64 t[4] = {1, 0, 0, 0};
      for (int i = 0; i < 512; i++) {
        dbl_raw(t);
        if (i == 255) memcpy(one, t, 32);
      }
      memcpy(r2, t, 32);
    }
    static void dbl_raw(u64 t[4]) {
      u64 c = 0;
      for (int i = 0; i < 4; i++) {
        u64 n = (t[i] << 1) | c;
        c = t[i] >> 63;
        t[i] = n;
      }
      if (c || geq(t, P::MOD)) sub_raw(t, P::MOD);
    }
  };
  static const Consts& C() {
    static const Consts c;
    return c;
  }

  static bool geq(const u64 a[4], const u64 b[4]) {
    for (int i = 3; i >= 0; i--) {
      if (a[i] > b[i]) return true;
      if (a[i] < b[i]) return false;
    }
    return true;
  }
  static u64 sub_raw(u64 a[4], const u64 b[4]) {
    u64 br = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)a[i] - b[i] - br;
      a[i] = (u64)d;
      br = (u64)(d >> 64) & 1;
    }
    return br;
  }
  static u64 add_raw(u64 a[4], const u64 b[4]) {
    u64 c = 0;
    for (int i = 0; i < 4; i++) {
      u128 s = (u128)a[i] + b[i] + c;
      a[i] = (u64)s;
      c = (u64)(s >> 64);
    }
    return c;
  }

  static Fp zero() { return Fp{{0, 0, 0, 0}}; }
  static Fp one() {
    Fp r;
    memcpy(r.l, C().one, 32);
    return r;
  }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const Fp& o) const { return memcmp(l, o.l, 32) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  Fp operator+(const Fp& o) const {
    Fp r = *this;
    u64 c = add_raw(r.l, o.l);
    if (c || geq(r.l, P::MOD)) sub_raw(r.l, P::MOD);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r = *this;
    if (sub_raw(r.l, o.l)) add_raw(r.l, P::MOD);
    return r;
  }
  Fp neg() const { return zero() - *this; }
  Fp dbl() const { return *this + *this; }

  // Montgomery product (CIOS, 64-bit limbs)
  static void mont_mul(u64 out[4], const u64 a[4], const u64 b[4]) {
    const u64 inv = C().inv;
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u64 c = 0;
      for (int j = 0; j < 4; j++) {
        u128 s = (u128)a[j] * b[i] + t[j] + c;
        t[j] = (u64)s;
        c = (u64)(s >> 64);
      }
      u128 s = (u128)t[4] + c;
      t[4] = (u64)s;
      t[5] = (u64)(s >> 64);
      u64 m = t[0] * inv;
      s = (u128)m * P::MOD[0] + t[0];
      c = (u64)(s >> 64);
      for (int j = 1; j < 4; j++) {
        s = (u128)m * P::MOD[j] + t[j] + c;
        t[j - 1] = (u64)s;
        c = (u64)(s >> 64);
      }
      s = (u128)t[4] + c;
      t[3] = (u64)s;
      t[4] = t[5] + (u64)(s >> 64);
    }
    if (t[4] || geq(t, P::MOD)) sub_raw(t, P::MOD);
    memcpy(out, t, 32);
  }
  Fp operator*(const Fp& o) const {
    Fp r;
    mont_mul(r.l, l, o.l);
    return r;
  }
  Fp sqr() const { return *this * *this; }
  Fp& operator+=(const Fp& o) { return *this = *this + o; }
  Fp& operator-=(const Fp& o) { return *this = *this - o; }
  Fp& operator*=(const Fp& o) { return *this = *this * o; }

  // canonical (non-Montgomery) limbs <-> Montgomery
  static Fp from_canonical(const u64 c[4]) {  // c < p
    Fp r;
    mont_mul(r.l, c, C().r2);
    return r;
  }
  void to_canonical(u64 out[4]) const {
    const u64 one_raw[4] = {1, 0, 0, 0};
    mont_mul(out, l, one_raw);
  }
  static Fp from_u64(u64 v) {
    u64 c[4] = {v, 0, 0, 0};
    return from_canonical(c);
  }
  // 32-byte little-endian canonical serialization (ark-serialize for Fp)
  void to_bytes_le(uint8_t out[32]) const {
    u64 c[4];
    to_canonical(c);
    memcpy(out, c, 32);  // host is little-endian
  }
  // PrimeField::from_le_bytes_mod_order: int(LE bytes) mod p, any length
  static Fp from_le_bytes_mod_order(const uint8_t* b, size_t n) {
    // Horner over bytes from the most significant end: acc = acc*256 + byte
    Fp acc = zero();
    Fp f256 = from_u64(256);
    for (size_t i = n; i-- > 0;) acc = acc * f256 + from_u64(b[i]);
    return acc;
  }
  Fp pow(const u64 e[4]) const {
    Fp acc = one();
    for (int i = 255; i >= 0; i--) {
      acc = acc.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) acc = acc * *this;
    }
    return acc;
  }
  Fp inverse_fermat() const {  // a^(p-2); 0 -> 0.  Kept as the cross-check of inverse() (tests/test_oracle.py, field op 5)
    u64 e[4];
    memcpy(e, P::MOD, 32);
    e[0] -= 2;
    return pow(e);
  }
  // Field::inverse.  ark-ff 0.5.0 inverts with a binary extended Euclid on the Montgomery limbs (Guajardo, Kumar, Paar,
  // Pelzl, alg. 16) -- a few microseconds, not the 380 products of a Fermat power.  The restatement uses the same
  // family in Kaliski's form ("The Montgomery inverse and its applications", 1995), whose loop needs no modular
  // halving: phase 1 on plain integers (u, v, r, s) = (p, x, 0, 1) ends with r = -x^-1 2^k mod p, 254 <= k <= 508;
  // phase 2 removes 2^k and restores Montgomery form with two products.  The value is unique, so any algorithm gives
  // the reference's bytes; the choice only matters for the CPU baseline's per-commit normalisation leg
  // (kzg.rs:67-71), which a Fermat inverse made 4-6x slower than arkworks'.  0 -> 0.
  Fp inverse() const {
    if (is_zero()) return zero();
    u64 u[4], v[4], r[4] = {0, 0, 0, 0}, t[4] = {1, 0, 0, 0};
    memcpy(u, P::MOD, 32);
    memcpy(v, l, 32);  // x = aR as an integer < p
    auto shr1 = [](u64 x[4]) {
      x[0] = (x[0] >> 1) | (x[1] << 63);
      x[1] = (x[1] >> 1) | (x[2] << 63);
      x[2] = (x[2] >> 1) | (x[3] << 63);
      x[3] >>= 1;
    };
    auto shl1 = [](u64 x[4]) {
      x[3] = (x[3] << 1) | (x[2] >> 63);
      x[2] = (x[2] << 1) | (x[1] >> 63);
      x[1] = (x[1] << 1) | (x[0] >> 63);
      x[0] <<= 1;
    };
    int k = 0;
    while (v[0] | v[1] | v[2] | v[3]) {
      if (!(u[0] & 1)) {
        shr1(u);
        shl1(t);
      } else if (!(v[0] & 1)) {
        shr1(v);
        shl1(r);
      } else if (!geq(v, u)) {  // u > v
        sub_raw(u, v);
        shr1(u);
        add_raw(r, t);
        shl1(t);
      } else {
        sub_raw(v, u);
        shr1(v);
        add_raw(t, r);
        shl1(r);
      }
      k++;
    }
    if (geq(r, P::MOD)) sub_raw(r, P::MOD);
    u64 neg[4];
    memcpy(neg, P::MOD, 32);
    sub_raw(neg, r);  // x^-1 2^k mod p
    // (aR)^-1 2^k -> a^-1 R = (aR)^-1 R^2: multiply by 2^(512 - k) = R2 * 2^e / R^2 with e = 512 - k in [4, 258]
    int e = 512 - k, extra = 0;
    if (e > 255) {
      extra = e - 255;
      e = 255;
    }
    u64 pw[4] = {0, 0, 0, 0};
    pw[e / 64] = (u64)1 << (e % 64);  // any 256-bit value may be the second operand of the CIOS product
    Fp out;
    mont_mul(out.l, neg, C().r2);
    mont_mul(out.l, out.l, pw);
    for (int i = 0; i < extra; i++) out = out.dbl();
    return out;
  }
};

typedef Fp<FrTag> Fr;
typedef Fp<FqTag> Fq;

}  // namespace orc
