// CPU ORACLE (test infrastructure, NOT product code).  PARITY UNPINNED -- see oracle/README.md.
// extern "C" surface of the oracle for ctypes (tests/, __graft_entry__.smoke(), bench.py cpu_baseline only).
// Conventions mirror the product C ABI (include/quill_b200.h): Fr/Fq values are 32-byte little-endian
// Montgomery limbs (the in-memory layout of ark_bn254::Fr), G1 affine points are x‖y (64 B, Montgomery) with
// (0,0) standing for the point at infinity, transcript state is the 32-byte blake3 state in/out.
#include <chrono>
#include "pairing.hpp"
#include "protocol.hpp"

using namespace orc;

static Fr ld_fr(const uint8_t* p) {
  Fr r;
  memcpy(r.l, p, 32);
  return r;
}
static void st_fr(uint8_t* p, const Fr& v) { memcpy(p, v.l, 32); }
static G1Affine ld_aff(const uint8_t* p) {
  G1Affine a;
  memcpy(a.x.l, p, 32);
  memcpy(a.y.l, p + 32, 32);
  a.inf = a.x.is_zero() && a.y.is_zero();
  return a;
}
static void st_aff(uint8_t* p, const G1Affine& a) {
  if (a.inf) {
    memset(p, 0, 64);
    return;
  }
  memcpy(p, a.x.l, 32);
  memcpy(p + 32, a.y.l, 32);
}
static Expr ld_expr(const uint32_t* nodes, size_t n_nodes, const uint8_t* consts, size_t n_consts) {
  Expr e;
  for (size_t i = 0; i < n_nodes; i++) e.nodes.push_back(ExprNode{nodes[3 * i], nodes[3 * i + 1], nodes[3 * i + 2]});
  for (size_t i = 0; i < n_consts; i++) e.consts.push_back(ld_fr(consts + 32 * i));
  return e;
}

extern "C" {

// ---- hashing / transcript -------------------------------------------------------------------------------------
void orc_blake3(const uint8_t* in, size_t n, uint8_t* out, size_t out_len) {
  Blake3 h;
  h.update(in, n);
  h.finalize(out, out_len);
}
void orc_transcript_new(const uint8_t* domain, size_t n, uint8_t state[32]) {
  Transcript t(domain, n);
  memcpy(state, t.state, 32);
}
void orc_transcript_append(uint8_t state[32], const uint8_t* msg, size_t n) {
  Transcript t(state);
  t.append_bytes(msg, n);
  memcpy(state, t.state, 32);
}
void orc_transcript_draw_fr(uint8_t state[32], uint8_t out_mont[32]) {
  Transcript t(state);
  st_fr(out_mont, t.draw_field_element());
  memcpy(state, t.state, 32);
}

// ---- field helpers (field: 0 = Fr, 1 = Fq); op: 0 add, 1 sub, 2 mul, 3 inverse(a), 4 to_mont(a canonical), 5 from_mont,
// 6 inverse by Fermat (the cross-check of 3)
void orc_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (field == 0) {
      Fr x = ld_fr(a + 32 * i), y = b ? ld_fr(b + 32 * i) : Fr::zero(), r;
      switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x.inverse(); break;
        case 4: r = Fr::from_canonical(x.l); break;
        case 6: r = x.inverse_fermat(); break;
        default: x.to_canonical(r.l); break;
      }
      memcpy(out + 32 * i, r.l, 32);
    } else {
      Fq x, y = Fq::zero(), r;
      memcpy(x.l, a + 32 * i, 32);
      if (b) memcpy(y.l, b + 32 * i, 32);
      switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x.inverse(); break;
        case 4: r = Fq::from_canonical(x.l); break;
        case 6: r = x.inverse_fermat(); break;
        default: x.to_canonical(r.l); break;
      }
      memcpy(out + 32 * i, r.l, 32);
    }
  }
}
void orc_fr_from_le_bytes_mod_order(const uint8_t* b, size_t n, uint8_t out_mont[32]) {
  st_fr(out_mont, Fr::from_le_bytes_mod_order(b, n));
}

// ---- G1 -------------------------------------------------------------------------------------------------------
int orc_g1_on_curve(const uint8_t* xy) { return ld_aff(xy).on_curve() ? 1 : 0; }
void orc_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  st_aff(out, G1::from_affine(ld_aff(a)).add(G1::from_affine(ld_aff(b))).into_affine());
}
void orc_g1_mul(const uint8_t* a, const uint8_t* scalar_mont, uint8_t* out) {
  st_aff(out, G1::from_affine(ld_aff(a)).mul(ld_fr(scalar_mont)).into_affine());
}
void orc_g1_serialize(const uint8_t* xy, uint8_t out[64]) { g1_serialize_uncompressed(ld_aff(xy), out); }

// SRS generator following pcs/src/kzg.rs:35-59: points[i] = g * tau^i, i < n, returned affine.  Uses a fixed-base
// 8-bit window table instead of n full double-and-add multiplications (same points, faster).
void orc_srs_generate(const uint8_t* g_xy, const uint8_t* tau_mont, size_t n, uint8_t* out_xy, int threads) {
  G1 g = G1::from_affine(ld_aff(g_xy));
  Fr tau = ld_fr(tau_mont);
  // table[w][d-1] = (d << 8w) * g
  std::vector<G1Affine> table(32 * 255);
  {
    std::vector<G1> tj(32 * 255);
    G1 base = g;
    for (int w = 0; w < 32; w++) {
      G1 acc = base;
      for (int d = 1; d <= 255; d++) {
        tj[w * 255 + d - 1] = acc;
        acc = acc.add(base);
      }
      base = acc;  // 256 * base
    }
    for (size_t i = 0; i < tj.size(); i++) table[i] = tj[i].into_affine();
  }
  if (threads < 1) threads = 1;
  auto work = [&](int t) {
    size_t lo = n * t / threads, hi = n * (t + 1) / threads;
    if (lo >= hi) return;
    u64 e[4] = {lo, 0, 0, 0};
    Fr ti = tau.pow(e);
    for (size_t i = lo; i < hi; i++) {
      u64 k[4];
      ti.to_canonical(k);
      G1 acc = G1::identity();
      for (int w = 0; w < 32; w++) {
        unsigned d = (unsigned)((k[w / 8] >> (8 * (w % 8))) & 0xff);
        if (d) acc = acc.add_affine(table[w * 255 + d - 1]);
      }
      st_aff(out_xy + 64 * i, acc.into_affine());
      ti *= tau;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(work, t);
  for (auto& th : pool) th.join();
}

// MSM with msm_unchecked semantics (pcs/src/kzg.rs:72): zip to the shorter.  mode 0 = naive, 1 = Pippenger.
void orc_msm(const uint8_t* bases_xy, size_t n_bases, const uint8_t* scalars_mont, size_t n_scalars, int mode,
             int threads, uint8_t out_xy[64]) {
  size_t n = std::min(n_bases, n_scalars);
  std::vector<G1Affine> b(n);
  std::vector<Fr> s(n);
  for (size_t i = 0; i < n; i++) {
    b[i] = ld_aff(bases_xy + 64 * i);
    s[i] = ld_fr(scalars_mont + 32 * i);
  }
  G1 r = mode == 0 ? msm_naive(b.data(), s.data(), n) : msm_pippenger(b.data(), s.data(), n, threads);
  st_aff(out_xy, r.into_affine());
}

// KZG::commit exactly as pcs/src/kzg.rs:61-73 does it: the SRS is held in projective form and every point is
// re-normalised (one Fq inversion each) on every call, then Pippenger.  The SRS here is given affine and lifted to
// Jacobian with a non-trivial Z (Z = z_seed^(i+1)) so the per-point inversion is real work, as in the reference.
// Returns seconds spent (normalise, msm) in out_secs[2].  rc: 0 ok, 1 = polynomial longer than the SRS (reference panics).
int orc_kzg_commit_reference_shape(const uint8_t* bases_xy, size_t n_bases, const uint8_t* scalars_mont,
                                   size_t n_scalars, int threads, uint8_t out_xy[64], double out_secs[2]) {
  if (n_scalars > n_bases) return 1;
  std::vector<G1> jac(n_bases);
  Fq z = Fq::from_u64(7);
  for (size_t i = 0; i < n_bases; i++) {
    G1Affine a = ld_aff(bases_xy + 64 * i);
    if (a.inf) {
      jac[i] = G1::identity();
    } else {
      Fq z2 = z.sqr();
      jac[i] = G1{a.x * z2, a.y * z2 * z, z};
    }
    z = z * Fq::from_u64(7) + Fq::one();
    if (z.is_zero()) z = Fq::one();
  }
  auto t0 = std::chrono::steady_clock::now();
  std::vector<G1Affine> aff(n_bases);
  for (size_t i = 0; i < n_bases; i++) aff[i] = jac[i].into_affine();
  auto t1 = std::chrono::steady_clock::now();
  std::vector<Fr> s(n_scalars);
  for (size_t i = 0; i < n_scalars; i++) s[i] = ld_fr(scalars_mont + 32 * i);
  G1 r = msm_pippenger(aff.data(), s.data(), std::min(n_bases, n_scalars), threads);
  auto t2 = std::chrono::steady_clock::now();
  st_aff(out_xy, r.into_affine());
  if (out_secs) {
    out_secs[0] = std::chrono::duration<double>(t1 - t0).count();
    out_secs[1] = std::chrono::duration<double>(t2 - t1).count();
  }
  return 0;
}

// KZG::open (pcs/src/kzg.rs:75-96): y, quotient coefficients (len-1 slots, zero padded), quotient length.
void orc_kzg_open_quotient(const uint8_t* poly_mont, size_t len, const uint8_t* x_mont, uint8_t* y_mont,
                           uint8_t* q_mont, size_t* q_len) {
  std::vector<Fr> p(len), q;
  for (size_t i = 0; i < len; i++) p[i] = ld_fr(poly_mont + 32 * i);
  Fr y;
  kzg_open_quotient(p.data(), len, ld_fr(x_mont), y, q);
  st_fr(y_mont, y);
  for (size_t i = 0; i < q.size(); i++) st_fr(q_mont + 32 * i, q[i]);
  *q_len = q.size();
}

// ---- eq table ---------------------------------------------------------------------------------------------------
void orc_eq_table(size_t n, const uint8_t* point_mont, uint8_t* out_mont) {
  std::vector<Fr> pt(n);
  for (size_t i = 0; i < n; i++) pt[i] = ld_fr(point_mont + 32 * i);
  std::vector<Fr> t = fast_eq_eval_hypercube(n, pt.data());
  memcpy(out_mont, t.data(), 32 * t.size());
}
void orc_eq_eval(size_t n, const uint8_t* x_mont, const uint8_t* r_mont, uint8_t* out_mont) {
  std::vector<Fr> x(n), r(n);
  for (size_t i = 0; i < n; i++) {
    x[i] = ld_fr(x_mont + 32 * i);
    r[i] = ld_fr(r_mont + 32 * i);
  }
  st_fr(out_mont, eq_eval(x.data(), r.data(), n));
}
// DenseMultilinearExtension::evaluate (variable j <-> index bit j), used by the reference's tests as the check
void orc_mle_evaluate(size_t n, const uint8_t* evals_mont, const uint8_t* point_mont, uint8_t* out_mont) {
  std::vector<Fr> t((size_t)1 << n);
  memcpy((void*)t.data(), evals_mont, 32 * t.size());
  for (size_t j = 0; j < n; j++) {
    Fr r = ld_fr(point_mont + 32 * j);
    size_t half = t.size() >> 1;
    for (size_t p = 0; p < half; p++) t[p] = t[2 * p] + r * (t[2 * p + 1] - t[2 * p]);
    t.resize(half);
  }
  st_fr(out_mont, t[0]);
}

// ---- sumcheck / zerocheck ------------------------------------------------------------------------------------------
// out_coeffs: num_vars * max_coeffs * 32 B (Montgomery, zero padded); out_lens: num_vars; out_point: num_vars * 32 B.
// zerocheck != 0 runs ZeroCheckProof::prove (claimed_sum ignored, out_z gets the num_vars eq challenges).
// rc: 0 ok, 2 = degree does not fit max_coeffs.
int orc_sumcheck_prove(size_t num_vars, size_t k, const uint8_t* const* tables_mont, const uint32_t* nodes,
                       size_t n_nodes, const uint8_t* consts_mont, size_t n_consts, const uint8_t* claimed_sum_mont,
                       uint8_t state[32], size_t max_coeffs, uint8_t* out_coeffs, uint32_t* out_lens,
                       uint8_t* out_point, uint8_t* out_eval, int zerocheck, uint8_t* out_z, int threads) {
  std::vector<const Fr*> tabs(k);
  for (size_t i = 0; i < k; i++) tabs[i] = (const Fr*)tables_mont[i];
  Expr h = ld_expr(nodes, n_nodes, consts_mont, n_consts);
  Transcript tr(state);
  SumcheckOutput sc;
  Fr ev;
  try {
    if (zerocheck) {
      ZerocheckOutput z = zerocheck_prove(num_vars, tabs, h, tr, threads);
      sc = z.sc;
      ev = z.evaluation;
      for (size_t i = 0; i < num_vars; i++) st_fr(out_z + 32 * i, z.z[i]);
    } else {
      sc = sumcheck_prove(num_vars, tabs, h, ld_fr(claimed_sum_mont), tr, threads);
      ev = sc.evaluation;
    }
  } catch (const std::exception&) {
    return 2;
  }
  memset(out_coeffs, 0, num_vars * max_coeffs * 32);
  for (size_t j = 0; j < num_vars; j++) {
    if ((size_t)sc.r_polys[j].n > max_coeffs) return 2;
    out_lens[j] = (uint32_t)sc.r_polys[j].n;
    for (int c = 0; c < sc.r_polys[j].n; c++) st_fr(out_coeffs + (j * max_coeffs + c) * 32, sc.r_polys[j].c[c]);
    st_fr(out_point + 32 * j, sc.point[j]);
  }
  st_fr(out_eval, ev);
  memcpy(state, tr.state, 32);
  return 0;
}

// SumcheckProof::verify (sumcheck.rs:116-150).  rc 0 = accepted (point/eval written), 1 = rejected.
int orc_sumcheck_verify(size_t num_vars, const uint8_t* claimed_sum_mont, size_t max_coeffs, const uint8_t* coeffs,
                        const uint32_t* lens, uint8_t state[32], uint8_t* out_point, uint8_t* out_eval) {
  std::vector<Poly> polys(num_vars);
  for (size_t j = 0; j < num_vars; j++) {
    polys[j].n = (int)lens[j];
    for (uint32_t c = 0; c < lens[j]; c++) polys[j].c[c] = ld_fr(coeffs + (j * max_coeffs + c) * 32);
  }
  Transcript tr(state);
  std::vector<Fr> point;
  Fr ev;
  bool ok = sumcheck_verify(num_vars, ld_fr(claimed_sum_mont), polys, tr, point, ev);
  if (!ok) return 1;
  for (size_t j = 0; j < num_vars; j++) st_fr(out_point + 32 * j, point[j]);
  st_fr(out_eval, ev);
  memcpy(state, tr.state, 32);
  return 0;
}

// h(g_1, .., g_k) at a point (virtual_polynomial.rs:22-37, 323-331)
void orc_expr_eval_point(const uint32_t* nodes, size_t n_nodes, const uint8_t* consts_mont, size_t n_consts,
                         const uint8_t* g_mont, size_t k, uint8_t* out_mont) {
  Expr h = ld_expr(nodes, n_nodes, consts_mont, n_consts);
  std::vector<Fr> g(k);
  for (size_t i = 0; i < k; i++) g[i] = ld_fr(g_mont + 32 * i);
  st_fr(out_mont, h.eval_point(g.data()));
}


// InnerProductProof::compute_s_polynomial (pcs/src/ipa.rs:122-157); out has room for max(n1, n2) - 1 elements
void orc_compute_s_polynomial(const uint8_t* p1, size_t n1, const uint8_t* p2, size_t n2, uint8_t* out, size_t* out_len) {
  std::vector<Fr> a(n1), b(n2);
  for (size_t i = 0; i < n1; i++) a[i] = ld_fr(p1 + 32 * i);
  for (size_t i = 0; i < n2; i++) b[i] = ld_fr(p2 + 32 * i);
  std::vector<Fr> sp = compute_s_polynomial(a, b);
  for (size_t i = 0; i < sp.size(); i++) st_fr(out + 32 * i, sp[i]);
  *out_len = sp.size();
}
// P_r coefficients (pcs/src/mlpcs.rs:68-78), trimmed; out has room for 2^n elements
void orc_compute_pr(const uint8_t* point, size_t n, uint8_t* out, size_t* out_len) {
  std::vector<Fr> pt(n);
  for (size_t i = 0; i < n; i++) pt[i] = ld_fr(point + 32 * i);
  std::vector<Fr> pr = compute_pr(pt.data(), n);
  for (size_t i = 0; i < pr.size(); i++) st_fr(out + 32 * i, pr[i]);
  *out_len = pr.size();
}
// MLEvalProof::prove (pcs/src/mlpcs.rs:83-124) on an affine SRS.  out: evaluation (32) ‖ s_comm (64) ‖ 4 x [x (32) ‖ y (32)
// ‖ proof (64)] in the order poly_opening, poly_opening_inv, s_opening, s_opening_inv = 608 bytes.  rc 1 = degree too large.
int orc_mlpcs_open(const uint8_t* bases_xy, size_t n_bases, const uint8_t* poly, size_t len, const uint8_t* point,
                   size_t n_point, uint8_t state[32], int threads, uint8_t* out) {
  std::vector<G1Affine> srs(n_bases);
  for (size_t i = 0; i < n_bases; i++) srs[i] = ld_aff(bases_xy + 64 * i);
  std::vector<Fr> pt(n_point);
  for (size_t i = 0; i < n_point; i++) pt[i] = ld_fr(point + 32 * i);
  Transcript tr(state);
  MlEvalProof pf;
  try {
    pf = mlpcs_open(srs, (const Fr*)poly, len, pt.data(), n_point, tr, threads);
  } catch (const std::exception&) {
    return 1;
  }
  st_fr(out, pf.evaluation);
  st_aff(out + 32, pf.s_comm);
  const KzgOpening* ops[4] = {&pf.poly_opening, &pf.poly_opening_inv, &pf.s_opening, &pf.s_opening_inv};
  for (int i = 0; i < 4; i++) {
    uint8_t* o = out + 96 + 128 * i;
    st_fr(o, ops[i]->x);
    st_fr(o + 32, ops[i]->y);
    st_aff(o + 64, ops[i]->proof);
  }
  memcpy(state, tr.state, 32);
  return 0;
}


// Logup denominators (hyperplonk/src/piops/multiset_check.rs:43-95): out[i] = m(row_i) / (gamma + h(row_i)); n_nodes_m = 0
// means Equality mode (no multiplicities).  rc 1 = a denominator is zero (the reference panics on unwrap).
int orc_logup_denominators(size_t num_vars, size_t k, const uint8_t* const* tables_mont, const uint32_t* nodes_h,
                           size_t n_nodes_h, const uint32_t* nodes_m, size_t n_nodes_m, const uint8_t* consts_mont,
                           size_t n_consts, const uint8_t* gamma_mont, uint8_t* out_mont) {
  Expr h = ld_expr(nodes_h, n_nodes_h, consts_mont, n_consts);
  Expr m = ld_expr(nodes_m, n_nodes_m, consts_mont, n_consts);
  Fr gamma = ld_fr(gamma_mont);
  std::vector<Fr> g(k);
  for (size_t i = 0; i < ((size_t)1 << num_vars); i++) {
    for (size_t t = 0; t < k; t++) g[t] = ld_fr(tables_mont[t] + 32 * i);  // :46-50
    Fr d = gamma + h.eval_point(g.data());                                   // :51
    if (d.is_zero()) return 1;
    Fr v = d.inverse();
    if (n_nodes_m) v *= m.eval_point(g.data());                              // :85-87
    st_fr(out_mont + 32 * i, v);
  }
  return 0;
}

// ---- verifier side: G2 and the pairing (pcs/src/kzg.rs:49-52, 98-108) ---------------------------------------
// G2 affine = x.c0 ‖ x.c1 ‖ y.c0 ‖ y.c1 (128 B, Montgomery Fq), all-zero = infinity.
static G2Affine ld_g2(const uint8_t* p) {
  G2Affine a;
  memcpy(a.x.c0.l, p, 32);
  memcpy(a.x.c1.l, p + 32, 32);
  memcpy(a.y.c0.l, p + 64, 32);
  memcpy(a.y.c1.l, p + 96, 32);
  a.inf = a.x.is_zero() && a.y.is_zero();
  return a;
}
static void st_g2(uint8_t* p, const G2Affine& a) {
  if (a.inf) {
    memset(p, 0, 128);
    return;
  }
  memcpy(p, a.x.c0.l, 32);
  memcpy(p + 32, a.x.c1.l, 32);
  memcpy(p + 64, a.y.c0.l, 32);
  memcpy(p + 96, a.y.c1.l, 32);
}
void orc_g2_generator(uint8_t out[128]) { st_g2(out, g2_generator()); }
int orc_g2_on_curve(const uint8_t* a) { return ld_g2(a).on_curve() ? 1 : 0; }
void orc_g2_add(const uint8_t* a, const uint8_t* b, uint8_t* out) { st_g2(out, ld_g2(a).add(ld_g2(b))); }
void orc_g2_neg(const uint8_t* a, uint8_t* out) { st_g2(out, ld_g2(a).neg()); }
void orc_g2_mul(const uint8_t* a, const uint8_t* scalar_mont, uint8_t* out) {
  st_g2(out, ld_g2(a).mul(ld_fr(scalar_mont)));
}
// out = prod_i e(P_i, Q_i) as 12 Montgomery Fq coefficients; returns 1 when the product is one
int orc_pairing_product(const uint8_t* g1s, const uint8_t* g2s, size_t n, const uint8_t* final_exp, size_t exp_len,
                        uint8_t* out_fq12) {
  std::vector<G1Affine> ps(n);
  std::vector<G2Affine> qs(n);
  for (size_t i = 0; i < n; i++) {
    ps[i] = ld_aff(g1s + 64 * i);
    qs[i] = ld_g2(g2s + 128 * i);
  }
  Fq12 f = pairing_product(ps.data(), qs.data(), n, final_exp, exp_len);
  if (out_fq12)
    for (int i = 0; i < 12; i++) memcpy(out_fq12 + 32 * i, f.c[i].l, 32);
  return f == Fq12::one() ? 1 : 0;
}

}  // extern "C"
