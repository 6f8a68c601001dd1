// CPU ORACLE (test infrastructure, NOT product code).  PARITY UNPINNED -- see oracle/README.md.
// BN254 G1 (y^2 = x^3 + 3) restating what ark-ec 0.5.0 / ark-bn254 0.5.0 (Cargo.lock:24-25,36-37) give the
// reference at pcs/src/kzg.rs:61-73: `into_affine`, `VariableBaseMSM::msm_unchecked`, and the uncompressed
// serialization of a G1 point that feeds the transcript (pcs/src/mlpcs.rs:102).
#pragma once
#include <algorithm>
#include <thread>
#include "field.hpp"

namespace orc {

struct G1Affine {
  Fq x, y;
  bool inf;
  static G1Affine infinity() { return G1Affine{Fq::zero(), Fq::zero(), true}; }
  bool on_curve() const {
    if (inf) return true;
    return y.sqr() == x.sqr() * x + Fq::from_u64(3);
  }
  G1Affine neg() const { return G1Affine{x, y.neg(), inf}; }
};

// Jacobian coordinates (X/Z^2, Y/Z^3); Z = 0 is the identity.
struct G1 {
  Fq X, Y, Z;
  static G1 identity() { return G1{Fq::one(), Fq::one(), Fq::zero()}; }
  static G1 from_affine(const G1Affine& a) {
    if (a.inf) return identity();
    return G1{a.x, a.y, Fq::one()};
  }
  bool is_identity() const { return Z.is_zero(); }

  G1 dbl() const {  // dbl-2009-l (a = 0)
    if (is_identity()) return *this;
    Fq A = X.sqr(), B = Y.sqr(), C = B.sqr();
    Fq D = ((X + B).sqr() - A - C).dbl();
    Fq E = A.dbl() + A, F = E.sqr();
    G1 r;
    r.X = F - D.dbl();
    r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
    r.Z = (Y * Z).dbl();
    return r;
  }
  G1 add(const G1& o) const {  // add-2007-bl, complete via explicit branches
    if (is_identity()) return o;
    if (o.is_identity()) return *this;
    Fq Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
    Fq U1 = X * Z2Z2, U2 = o.X * Z1Z1;
    Fq S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
    if (U1 == U2) {
      if (S1 == S2) return dbl();
      return identity();
    }
    Fq H = U2 - U1, I = H.dbl().sqr(), J = H * I, rr = (S2 - S1).dbl(), V = U1 * I;
    G1 r;
    r.X = rr.sqr() - J - V.dbl();
    r.Y = rr * (V - r.X) - (S1 * J).dbl();
    r.Z = ((Z + o.Z).sqr() - Z1Z1 - Z2Z2) * H;
    return r;
  }
  G1 add_affine(const G1Affine& a) const {  // madd-2007-bl, complete via explicit branches
    if (a.inf) return *this;
    if (is_identity()) return from_affine(a);
    Fq Z1Z1 = Z.sqr();
    Fq U2 = a.x * Z1Z1, S2 = a.y * Z * Z1Z1;
    if (U2 == X) {
      if (S2 == Y) return dbl();
      return identity();
    }
    Fq H = U2 - X, HH = H.sqr(), I = HH.dbl().dbl(), J = H * I, rr = (S2 - Y).dbl(), V = X * I;
    G1 r;
    r.X = rr.sqr() - J - V.dbl();
    r.Y = rr * (V - r.X) - (Y * J).dbl();
    r.Z = (Z + H).sqr() - Z1Z1 - HH;
    return r;
  }
  G1 neg() const { return G1{X, Y.neg(), Z}; }

  // CurveGroup::into_affine (one Fq inversion; pcs/src/kzg.rs:67-71 does this per SRS point per commit)
  G1Affine into_affine() const {
    if (is_identity()) return G1Affine::infinity();
    Fq zi = Z.inverse(), zi2 = zi.sqr();
    return G1Affine{X * zi2, Y * zi2 * zi, false};
  }
  // scalar given as canonical 4x64 limbs
  G1 mul_canonical(const u64 k[4]) const {
    G1 acc = identity();
    for (int i = 255; i >= 0; i--) {
      acc = acc.dbl();
      if ((k[i / 64] >> (i % 64)) & 1) acc = acc.add(*this);
    }
    return acc;
  }
  G1 mul(const Fr& s) const {
    u64 k[4];
    s.to_canonical(k);
    return mul_canonical(k);
  }
};

// ark-serialize uncompressed G1: x (32 B LE) ‖ y (32 B LE); flags in the top bits of the last byte:
// 0x80 when y > -y (y > (q-1)/2), 0x40 with zero coordinates for the point at infinity.
inline void g1_serialize_uncompressed(const G1Affine& a, uint8_t out[64]) {
  if (a.inf) {
    memset(out, 0, 64);
    out[63] |= 0x40;
    return;
  }
  a.x.to_bytes_le(out);
  a.y.to_bytes_le(out + 32);
  u64 yc[4], nc[4];
  a.y.to_canonical(yc);
  a.y.neg().to_canonical(nc);
  bool y_gt_neg = Fq::geq(yc, nc) && memcmp(yc, nc, 32) != 0;
  if (y_gt_neg) out[63] |= 0x80;
}

// ---- MSM ------------------------------------------------------------------------------------------------
inline G1 msm_naive(const G1Affine* bases, const Fr* scalars, size_t n) {
  G1 acc = G1::identity();
  for (size_t i = 0; i < n; i++) acc = acc.add(G1::from_affine(bases[i]).mul(scalars[i]));
  return acc;
}

// ark-ec's window heuristic (recalled; affects only CPU timing, never the result)
inline int ark_window_bits(size_t n) {
  if (n < 32) return 3;
  int lg = 63 - __builtin_clzll((unsigned long long)n);
  return lg * 69 / 100 + 2;
}

// Signed-digit windowed Pippenger in the shape of ark-ec 0.5.0 `msm_bigint_wnaf`: per-window bucket arrays of
// 2^(c-1) buckets, running-sum reduction, c doublings between windows (highest window first).
// `threads` > 1 spreads windows over std::threads (generous to the reference, which has no `parallel`).
inline G1 msm_pippenger(const G1Affine* bases, const Fr* scalars, size_t n, int threads = 1, int c_override = 0) {
  if (n == 0) return G1::identity();
  const int c = c_override ? c_override : ark_window_bits(n);
  const int num_bits = 254;
  const int W = (num_bits + c - 1) / c + 1;  // +1 window absorbs the final signed carry
  // signed digits, digit-major per scalar
  std::vector<int32_t> digits(n * (size_t)W);
  for (size_t i = 0; i < n; i++) {
    u64 k[4];
    scalars[i].to_canonical(k);
    int64_t carry = 0;
    for (int w = 0; w < W; w++) {
      int bit = w * c;
      u64 v = 0;
      if (bit < 256) {
        int limb = bit / 64, off = bit % 64;
        v = k[limb] >> off;
        if (off + c > 64 && limb + 1 < 4) v |= k[limb + 1] << (64 - off);
        v &= ((u64)1 << c) - 1;
      }
      int64_t d = (int64_t)v + carry;
      carry = 0;
      if (d > ((int64_t)1 << (c - 1))) {
        d -= (int64_t)1 << c;
        carry = 1;
      }
      digits[i * W + w] = (int32_t)d;
    }
  }
  std::vector<G1> wsum(W, G1::identity());
  auto do_window = [&](int w) {
    std::vector<G1> buckets((size_t)1 << (c - 1), G1::identity());
    for (size_t i = 0; i < n; i++) {
      int32_t d = digits[i * W + w];
      if (d > 0)
        buckets[d - 1] = buckets[d - 1].add_affine(bases[i]);
      else if (d < 0)
        buckets[-d - 1] = buckets[-d - 1].add_affine(bases[i].neg());
    }
    G1 running = G1::identity(), acc = G1::identity();
    for (size_t b = buckets.size(); b-- > 0;) {
      running = running.add(buckets[b]);
      acc = acc.add(running);
    }
    wsum[w] = acc;
  };
  if (threads <= 1) {
    for (int w = 0; w < W; w++) do_window(w);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
      pool.emplace_back([&, t]() {
        for (int w = t; w < W; w += threads) do_window(w);
      });
    for (auto& th : pool) th.join();
  }
  G1 total = wsum[W - 1];
  for (int w = W - 2; w >= 0; w--) {
    for (int j = 0; j < c; j++) total = total.dbl();
    total = total.add(wsum[w]);
  }
  return total;
}

}  // namespace orc
