// CPU ORACLE (test infrastructure, NOT product code).  PARITY UNPINNED -- see oracle/README.md.
// Restates, on the CPU, the reference's transcript, expression tree, sumcheck / zero-check provers and
// verifiers, eq table and KZG commit/open.  Every function cites the reference file:line it follows.
#pragma once
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "blake3_ref.hpp"
#include "g1.hpp"

namespace orc {

// ---- transcript/src/transcript.rs:14-75 ------------------------------------------------------------------
struct Transcript {
  uint8_t state[32];
  explicit Transcript(const uint8_t* domain, size_t n) { Blake3::hash(domain, n, state); }  // :15-23
  explicit Transcript(const uint8_t st[32]) { memcpy(state, st, 32); }
  void append_bytes(const uint8_t* msg, size_t n) {  // :26-32
    Blake3 h;
    h.update(state, 32);
    h.update(msg, n);
    h.finalize(state, 32);
  }
  void append_usize(u64 v) { append_bytes((const uint8_t*)&v, 8); }  // ark-serialize: usize -> u64 LE
  void append_fr(const Fr& v) {                                      // Fr -> 32 B LE canonical
    uint8_t b[32];
    v.to_bytes_le(b);
    append_bytes(b, 32);
  }
  void append_fr_vec(const Fr* v, size_t n) {  // Vec / slice / DensePolynomial: u64 LE len ‖ elements
    std::vector<uint8_t> b(8 + 32 * n);
    u64 len = n;
    memcpy(b.data(), &len, 8);
    for (size_t i = 0; i < n; i++) v[i].to_bytes_le(b.data() + 8 + 32 * i);
    append_bytes(b.data(), b.size());
  }
  void append_g1(const G1Affine& p) {
    uint8_t b[64];
    g1_serialize_uncompressed(p, b);
    append_bytes(b, 64);
  }
  void draw_challenge(uint8_t* out, size_t n) {  // :49-63
    Blake3 h;
    h.update(state, 32);
    h.update((const uint8_t*)"challenge", 9);
    h.finalize(out, n);
    append_bytes(out, n);
  }
  Fr draw_field_element() {  // :71-75; (254 + 128 + 7) / 8 = 48 bytes
    uint8_t b[48];
    draw_challenge(b, 48);
    return Fr::from_le_bytes_mod_order(b, 48);
  }
};

// ---- small dense univariate polynomial (ark-poly DensePolynomial semantics: trailing zeros trimmed) ----------
constexpr int MAXC = 33;  // max coefficients (degree <= 32)
struct Poly {
  Fr c[MAXC];
  int n = 0;
  Poly() {}
  Poly(const Poly& o) : n(o.n) { memcpy((void*)c, (const void*)o.c, sizeof(Fr) * (size_t)o.n); }
  Poly& operator=(const Poly& o) {
    n = o.n;
    memcpy((void*)c, (const void*)o.c, sizeof(Fr) * (size_t)o.n);
    return *this;
  }
  void trim() {
    while (n > 0 && c[n - 1].is_zero()) n--;
  }
  static Poly constant(const Fr& v) {
    Poly p;
    p.c[0] = v;
    p.n = 1;
    p.trim();
    return p;
  }
  static Poly linear(const Fr& c0, const Fr& c1) {
    Poly p;
    p.c[0] = c0;
    p.c[1] = c1;
    p.n = 2;
    p.trim();
    return p;
  }
  Poly add(const Poly& o) const {
    Poly r;
    r.n = std::max(n, o.n);
    for (int i = 0; i < r.n; i++) r.c[i] = (i < n ? c[i] : Fr::zero()) + (i < o.n ? o.c[i] : Fr::zero());
    r.trim();
    return r;
  }
  Poly mul(const Poly& o) const {
    Poly r;
    if (n == 0 || o.n == 0) return r;
    if (n + o.n - 1 > MAXC) throw std::runtime_error("oracle: expression degree exceeds MAXC-1");
    r.n = n + o.n - 1;
    for (int i = 0; i < r.n; i++) r.c[i] = Fr::zero();
    for (int i = 0; i < n; i++)
      for (int j = 0; j < o.n; j++) r.c[i + j] += c[i] * o.c[j];
    r.trim();
    return r;
  }
  Fr eval(const Fr& x) const {
    Fr acc = Fr::zero();
    for (int i = n; i-- > 0;) acc = acc * x + c[i];
    return acc;
  }
};

// ---- hyperplonk/src/utils/virtual_polynomial.rs:9-18 -- flattened tree, children before parents, root last --
enum : uint32_t { EX_INPUT = 0, EX_CONST = 1, EX_ADD = 2, EX_MUL = 3 };
struct ExprNode {
  uint32_t op, a, b;
};
struct Expr {
  std::vector<ExprNode> nodes;
  std::vector<Fr> consts;
  Fr eval_point(const Fr* g, int node = -1) const {  // virtual_polynomial.rs:22-37
    if (node < 0) node = (int)nodes.size() - 1;
    const ExprNode& e = nodes[node];
    switch (e.op) {
      case EX_INPUT: return g[e.a];
      case EX_CONST: return consts[e.a];
      case EX_ADD: return eval_point(g, e.a) + eval_point(g, e.b);
      default: return eval_point(g, e.a) * eval_point(g, e.b);
    }
  }
  Poly eval_poly(const Poly* g, int node = -1) const {  // virtual_polynomial.rs:300-320
    if (node < 0) node = (int)nodes.size() - 1;
    const ExprNode& e = nodes[node];
    switch (e.op) {
      case EX_INPUT: return g[e.a];
      case EX_CONST: return Poly::constant(consts[e.a]);
      case EX_ADD: return eval_poly(g, e.a).add(eval_poly(g, e.b));
      default: return eval_poly(g, e.a).mul(eval_poly(g, e.b));
    }
  }
};

struct SumcheckOutput {
  std::vector<Poly> r_polys;
  std::vector<Fr> point;
  Fr evaluation = Fr::zero();
};

// ---- hyperplonk/src/piops/sumcheck.rs:28-114 ---------------------------------------------------------------
// `tables` are the store's polynomials (all of them are cloned and folded, :44-49).  threads > 1 splits the
// per-round loops over std::threads (the reference itself is single-threaded).
inline SumcheckOutput sumcheck_prove(size_t num_vars, const std::vector<const Fr*>& tables, const Expr& h,
                                     const Fr& claimed_sum, Transcript& tr, int threads = 1) {
  tr.append_usize(num_vars);   // :35
  tr.append_fr(claimed_sum);   // :36
  const size_t k = tables.size();
  std::vector<std::vector<Fr>> gs(k);
  for (size_t t = 0; t < k; t++) gs[t].assign(tables[t], tables[t] + ((size_t)1 << num_vars));
  SumcheckOutput out;
  for (size_t i = num_vars; i-- > 0;) {  // :51
    const size_t pairs = (size_t)1 << i;
    int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, pairs / 1024 + 1));
    std::vector<Poly> partial(T);
    auto eval_range = [&](int t) {
      size_t lo = pairs * t / T, hi = pairs * (t + 1) / T;
      Poly acc;
      std::vector<Poly> lin(k);
      for (size_t p = lo; p < hi; p++) {  // :53-63
        for (size_t g = 0; g < k; g++) lin[g] = Poly::linear(gs[g][2 * p], gs[g][2 * p + 1] - gs[g][2 * p]);
        acc = acc.add(h.eval_poly(lin.data()));  // :67-70
      }
      partial[t] = acc;
    };
    if (T == 1) {
      eval_range(0);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < T; t++) pool.emplace_back(eval_range, t);
      for (auto& th : pool) th.join();
    }
    Poly msg;
    for (int t = 0; t < T; t++) msg = msg.add(partial[t]);
    tr.append_fr_vec(msg.c, msg.n);  // :73
    out.r_polys.push_back(msg);
    Fr r = tr.draw_field_element();  // :77
    out.point.push_back(r);
    // :81-92: g'[p] = low + r*(high-low)
    for (size_t g = 0; g < k; g++) {
      std::vector<Fr> ng(pairs);
      auto fr = [&](int t) {
        size_t lo = pairs * t / T, hi = pairs * (t + 1) / T;
        for (size_t p = lo; p < hi; p++) ng[p] = gs[g][2 * p] + r * (gs[g][2 * p + 1] - gs[g][2 * p]);
      };
      if (T == 1) {
        fr(0);
      } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < T; t++) pool.emplace_back(fr, t);
        for (auto& th : pool) th.join();
      }
      gs[g].swap(ng);
    }
    if (i == 0) {  // :94-100
      std::vector<Fr> fin(k);
      for (size_t g = 0; g < k; g++) fin[g] = gs[g][0];
      out.evaluation = h.eval_point(fin.data());
    }
  }
  return out;
}

// ---- sumcheck.rs:116-150 -----------------------------------------------------------------------------------
inline bool sumcheck_verify(size_t num_vars, const Fr& claimed_sum, const std::vector<Poly>& r_polys,
                            Transcript& tr, std::vector<Fr>& point, Fr& evaluation) {
  tr.append_usize(num_vars);
  tr.append_fr(claimed_sum);
  Fr v = claimed_sum;
  point.clear();
  for (const Poly& p : r_polys) {
    if (p.eval(Fr::zero()) + p.eval(Fr::one()) != v) return false;
    tr.append_fr_vec(p.c, p.n);
    Fr r = tr.draw_field_element();
    point.push_back(r);
    v = p.eval(r);
  }
  evaluation = v;
  return true;
}

// ---- hyperplonk/src/utils/eq_eval.rs:6-31 ------------------------------------------------------------------
inline std::vector<Fr> fast_eq_eval_hypercube(size_t n, const Fr* point) {
  std::vector<Fr> evals{Fr::one()};
  for (size_t i = n; i-- > 0;) {
    Fr r = point[i], om = Fr::one() - r;
    std::vector<Fr> nw;
    nw.reserve(evals.size() * 2);
    for (const Fr& e : evals) {
      nw.push_back(e * om);
      nw.push_back(e * r);
    }
    evals.swap(nw);
  }
  return evals;
}
// ---- eq_eval.rs:33-43 ----
inline Fr eq_eval(const Fr* x, const Fr* r, size_t n) {
  Fr res = Fr::one();
  for (size_t i = 0; i < n; i++) res *= x[i] * r[i] + (Fr::one() - x[i]) * (Fr::one() - r[i]);
  return res;
}

// ---- hyperplonk/src/piops/zerocheck.rs:14-49 ---------------------------------------------------------------
struct ZerocheckOutput {
  SumcheckOutput sc;
  std::vector<Fr> z;
  Fr evaluation;
};
inline ZerocheckOutput zerocheck_prove(size_t num_vars, const std::vector<const Fr*>& tables, const Expr& h,
                                       Transcript& tr, int threads = 1) {
  ZerocheckOutput out;
  for (size_t i = 0; i < num_vars; i++) out.z.push_back(tr.draw_field_element());  // :20-22
  std::vector<Fr> eq = fast_eq_eval_hypercube(num_vars, out.z.data());              // :25
  std::vector<const Fr*> t2 = tables;
  t2.push_back(eq.data());  // :27 (eq table is appended as the last store polynomial)
  Expr hh = h;              // :28-29: h_hat = Mul(h, Input(eq))
  uint32_t root = (uint32_t)hh.nodes.size() - 1;
  hh.nodes.push_back(ExprNode{EX_INPUT, (uint32_t)tables.size(), 0});
  hh.nodes.push_back(ExprNode{EX_MUL, root, root + 1});
  out.sc = sumcheck_prove(num_vars, t2, hh, Fr::zero(), tr, threads);  // :31-32
  Fr e = eq_eval(out.z.data(), out.sc.point.data(), num_vars);         // :34
  out.evaluation = out.sc.evaluation * e.inverse();                     // :36-40
  return out;
}

// ---- pcs/src/kzg.rs:61-73 -- SRS kept in projective form like `KZG::g1_points` -------------------------------
inline G1 kzg_commit(const std::vector<G1>& g1_points, const Fr* poly, size_t len, int threads = 1,
                     bool naive = false) {
  if (len > g1_points.size()) throw std::runtime_error("Polynomial degree exceeds max degree");  // :62-65
  std::vector<G1Affine> aff(g1_points.size());
  for (size_t i = 0; i < aff.size(); i++) aff[i] = g1_points[i].into_affine();  // :67-71 (every call)
  size_t n = std::min(aff.size(), len);                                         // msm_unchecked: zip to shorter
  return naive ? msm_naive(aff.data(), poly, n) : msm_pippenger(aff.data(), poly, n, threads);  // :72
}

// ---- pcs/src/kzg.rs:75-96 (y and quotient only; the commit of q is the caller's) ----------------------------
inline void kzg_open_quotient(const Fr* poly, size_t len, const Fr& x, Fr& y, std::vector<Fr>& q) {
  while (len > 0 && poly[len - 1].is_zero()) len--;  // DensePolynomial::from_coefficients_slice trims
  y = Fr::zero();
  for (size_t i = len; i-- > 0;) y = y * x + poly[i];  // :78
  // (p - y) / (X - x): q_{i-1} = p_i + x*q_i (:81-84); the remainder is p(x) - y = 0
  q.assign(len > 0 ? len - 1 : 0, Fr::zero());
  Fr carry = Fr::zero();
  for (size_t i = len; i-- > 1;) {
    carry = poly[i] + carry * x;
    q[i - 1] = carry;
  }
  while (!q.empty() && q.back().is_zero()) q.pop_back();
}


// ---- radix-2 NTT over Fr (stands in for ark-poly's FFT-backed `&DensePolynomial * &DensePolynomial`) ---------
// The product of two polynomials does not depend on which primitive root of unity is used.
inline Fr root_of_unity(int log_n) {  // primitive 2^log_n-th root: 5^((r-1)/2^28) squared down; 5 generates Fr*
  static const Fr root28 = [] {
    u64 e[4];
    memcpy(e, FrTag::MOD, 32);
    e[0] -= 1;
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 28) | (i < 3 ? e[i + 1] << 36 : 0);
    return Fr::from_u64(5).pow(e);
  }();
  Fr w = root28;
  for (int i = 28; i > log_n; i--) w = w.sqr();
  return w;
}
inline void ntt(std::vector<Fr>& a, bool inverse) {
  const size_t n = a.size();
  int log_n = 0;
  while (((size_t)1 << log_n) < n) log_n++;
  for (size_t i = 1, j = 0; i < n; i++) {  // bit reversal
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (int s = 1; s <= log_n; s++) {
    Fr w = root_of_unity(s);
    if (inverse) w = w.inverse();
    const size_t half = (size_t)1 << (s - 1);
    std::vector<Fr> tw(half);
    tw[0] = Fr::one();
    for (size_t j = 1; j < half; j++) tw[j] = tw[j - 1] * w;
    for (size_t i = 0; i < n; i += 2 * half)
      for (size_t j = 0; j < half; j++) {
        Fr u = a[i + j], v = a[i + j + half] * tw[j];
        a[i + j] = u + v;
        a[i + j + half] = u - v;
      }
  }
  if (inverse) {
    Fr ninv = Fr::from_u64((u64)n).inverse();
    for (auto& x : a) x *= ninv;
  }
}
inline std::vector<Fr> poly_mul_ntt(const std::vector<Fr>& a, const std::vector<Fr>& b) {
  if (a.empty() || b.empty()) return {};
  size_t need = a.size() + b.size() - 1, m = 1;
  while (m < need) m <<= 1;
  std::vector<Fr> fa(a), fb(b);
  fa.resize(m, Fr::zero());
  fb.resize(m, Fr::zero());
  ntt(fa, false);
  ntt(fb, false);
  for (size_t i = 0; i < m; i++) fa[i] *= fb[i];
  ntt(fa, true);
  fa.resize(need);
  return fa;
}
inline void trim_vec(std::vector<Fr>& v) {
  while (!v.empty() && v.back().is_zero()) v.pop_back();
}

// ---- pcs/src/ipa.rs:122-157 ------------------------------------------------------------------------------------
inline std::vector<Fr> compute_s_polynomial(const std::vector<Fr>& p1, const std::vector<Fr>& p2) {
  const size_t L = std::max(p1.size(), p2.size());
  if (L == 0) return {};
  std::vector<Fr> f(p1), g(p2);
  f.resize(L, Fr::zero());
  g.resize(L, Fr::zero());
  std::vector<Fr> fr(f.rbegin(), f.rend()), gr(g.rbegin(), g.rend());
  // DensePolynomial::from_coefficients_* trims; products of trimmed polys, then h is re-padded to 2L-1 (:152)
  std::vector<Fr> ft(f), gt(g), frt(fr), grt(gr);
  trim_vec(ft), trim_vec(gt), trim_vec(frt), trim_vec(grt);
  std::vector<Fr> h1 = poly_mul_ntt(ft, grt), h2 = poly_mul_ntt(frt, gt);
  std::vector<Fr> h(2 * L - 1, Fr::zero());
  for (size_t i = 0; i < h1.size(); i++) h[i] += h1[i];
  for (size_t i = 0; i < h2.size(); i++) h[i] += h2[i];
  std::vector<Fr> sc(h.begin() + (h.size() / 2 + 1), h.end());  // :153
  trim_vec(sc);
  return sc;
}

// ---- pcs/src/mlpcs.rs:52-78: coefficients of P_r (= eq table of r, LSB-first), trailing zeros trimmed -------------
inline std::vector<Fr> compute_pr(const Fr* r, size_t n) {
  std::vector<Fr> pr = fast_eq_eval_hypercube(n, r);
  trim_vec(pr);
  return pr;
}

struct KzgOpening {
  Fr x, y;
  G1Affine proof;
};
struct MlEvalProof {
  Fr evaluation;
  G1Affine s_comm;
  KzgOpening poly_opening, poly_opening_inv, s_opening, s_opening_inv;
};
// KZG::open on an affine SRS (kzg.rs:75-96)
inline KzgOpening kzg_open(const std::vector<G1Affine>& srs, const Fr* poly, size_t len, const Fr& x, int threads) {
  KzgOpening o;
  o.x = x;
  std::vector<Fr> q;
  kzg_open_quotient(poly, len, x, o.y, q);
  if (q.size() > srs.size()) throw std::runtime_error("Polynomial degree exceeds max degree");
  o.proof = msm_pippenger(srs.data(), q.data(), q.size(), threads).into_affine();
  return o;
}
// ---- pcs/src/mlpcs.rs:83-124 -----------------------------------------------------------------------------------------
inline MlEvalProof mlpcs_open(const std::vector<G1Affine>& srs, const Fr* poly, size_t len, const Fr* point, size_t n,
                              Transcript& tr, int threads) {
  MlEvalProof pf;
  std::vector<Fr> pr = compute_pr(point, n);
  pf.evaluation = Fr::zero();
  for (size_t i = 0; i < std::min(len, pr.size()); i++) pf.evaluation += poly[i] * pr[i];  // :91-94
  std::vector<Fr> pv(poly, poly + len);
  std::vector<Fr> s = compute_s_polynomial(pv, pr);  // :96
  if (s.size() > srs.size() || len > srs.size()) throw std::runtime_error("Polynomial degree exceeds max degree");
  pf.s_comm = msm_pippenger(srs.data(), s.data(), s.size(), threads).into_affine();  // :97
  tr.append_fr_vec(point, n);       // :100 (&[F]: length-prefixed)
  tr.append_fr(pf.evaluation);      // :101
  tr.append_g1(pf.s_comm);          // :102
  Fr r = tr.draw_field_element();   // :105
  Fr r_inv = r.inverse();           // :107
  pf.poly_opening = kzg_open(srs, poly, len, r, threads);          // :109-113
  pf.poly_opening_inv = kzg_open(srs, poly, len, r_inv, threads);
  pf.s_opening = kzg_open(srs, s.data(), s.size(), r, threads);
  pf.s_opening_inv = kzg_open(srs, s.data(), s.size(), r_inv, threads);
  return pf;
}

}  // namespace orc
