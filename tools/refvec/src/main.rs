//! Reference vector emitter (see Cargo.toml).  Reads the INPUTS of tests/golden/golden.json (tables, scalars, points:
//! they come from Python's seeded `random`, which Rust cannot replay) and recomputes every OUTPUT with the reference's
//! own code: Transcript (transcript/src/transcript.rs), SumcheckProof::prove (hyperplonk/src/piops/sumcheck.rs:28),
//! ZeroCheckProof::prove (zerocheck.rs:14), fast_eq_eval_hypercube (utils/eq_eval.rs:6), KZG::commit / ::open
//! (pcs/src/kzg.rs:61,75), MLEvalProof::compute_pr / ::prove (pcs/src/mlpcs.rs:68,83),
//! InnerProductProof::compute_s_polynomial (pcs/src/ipa.rs:122) and HyperPlonk::prove (proof/proof.rs:239).
//! The output mirrors golden.json's layout key for key; tests/test_reference_vectors.py compares the two files.
//!
//! Encoding: a field element is 64 hex digits, big-endian, canonical (Python's "%064x"); a G1 point is [x, y] in that
//! form; `*_bytes` / `state*` fields are the raw ark-serialize / blake3 bytes in hex.
use ark_bn254::{Bn254, Fq, Fr, G1Affine, G1Projective, G2Projective};
use ark_ec::{AffineRepr, CurveGroup, PrimeGroup};
use ark_ff::{BigInteger, Field, PrimeField};
use ark_poly::univariate::DensePolynomial;
use ark_serialize::CanonicalSerialize;
use ark_std::{One, Zero};
use quill_hyperplonk::frontend::transition_circuit::TransitionCircuit;
use quill_hyperplonk::piops::sumcheck::SumcheckProof;
use quill_hyperplonk::piops::zerocheck::ZeroCheckProof;
use quill_hyperplonk::proof::circuit::Circuit;
use quill_hyperplonk::proof::proof::{HyperPlonk, TraceWitness};
use quill_hyperplonk::utils::eq_eval::fast_eq_eval_hypercube;
use quill_hyperplonk::utils::virtual_polynomial::{VirtualPolyExpr, VirtualPolynomialStore};
use quill_pcs::ipa::InnerProductProof;
use quill_pcs::kzg::KZG;
use quill_pcs::mlpcs::MLEvalProof;
use quill_transcript::transcript::Transcript;
use serde_json::{json, Map, Value};

fn hex(bytes: &[u8]) -> String {
    bytes.iter().map(|b| format!("{:02x}", b)).collect()
}
fn unhex(s: &str) -> Vec<u8> {
    (0..s.len() / 2).map(|i| u8::from_str_radix(&s[2 * i..2 * i + 2], 16).unwrap()).collect()
}
/// "%064x" of the canonical value
fn h<F: PrimeField>(x: &F) -> String {
    hex(&x.into_bigint().to_bytes_be())
}
fn fr(v: &Value) -> Fr {
    Fr::from_be_bytes_mod_order(&unhex(v.as_str().unwrap()))
}
fn frs(v: &Value) -> Vec<Fr> {
    v.as_array().unwrap().iter().map(fr).collect()
}
fn point(p: &G1Projective) -> Value {
    let a = p.into_affine();
    if a.is_zero() {
        return Value::Null;
    }
    json!([h(&a.x), h(&a.y)])
}
fn ser<T: CanonicalSerialize>(t: &T) -> String {
    let mut b = vec![];
    t.serialize_uncompressed(&mut b).unwrap();
    hex(&b)
}
fn polys(ps: &[DensePolynomial<Fr>]) -> Value {
    Value::Array(ps.iter().map(|p| Value::Array(p.coeffs.iter().map(|c| json!(h(c))).collect())).collect())
}
fn hs(xs: &[Fr]) -> Value {
    Value::Array(xs.iter().map(|x| json!(h(x))).collect())
}

/// the SRS make_golden.py builds: g = 7 * (1, 2), tau = 0x1234567890ABCDEF1234567890ABCDEF (kzg.rs:35-59 with the RNG
/// draws replaced by fixed values; every field of KZG is pub)
fn fixed_kzg(max_degree: usize) -> KZG<Bn254> {
    let g1 = G1Projective::generator() * Fr::from(7u64);
    let g2 = G2Projective::generator();
    let tau = Fr::from(0x1234567890ABCDEF1234567890ABCDEFu128);
    let mut g1_points = Vec::with_capacity(max_degree + 1);
    let mut t = Fr::one();
    for _ in 0..=max_degree {
        g1_points.push(g1 * t);
        t *= tau;
    }
    KZG { max_degree, g1, g2, g1_points, g2_points: vec![g2, g2 * tau] }
}

fn transcript_vectors() -> Value {
    let mut out = Map::new();
    let long = vec![b'x'; 100];
    let domains: [&[u8]; 5] = [b"", b"sumcheck_test", b"zerocheck_test", b"hyperplonk_proof", &long];
    for dom in domains {
        let mut t = Transcript::new(dom);
        let s0 = hex(&t.state);
        let f1: Fr = t.draw_field_element();
        t.append_serializable(&3usize);
        t.append_serializable(&Fr::from(48u64));
        t.append_serializable(&vec![Fr::from(0u64), Fr::from(38u64), Fr::from(10u64)]);
        let f2: Fr = t.draw_field_element();
        let c = t.draw_challenge(17);
        out.insert(
            String::from_utf8(dom.to_vec()).unwrap(),
            json!({"state0": s0, "fe1": h(&f1), "fe2": h(&f2), "challenge17": hex(&c), "state_end": hex(&t.state)}),
        );
    }
    Value::Object(out)
}

fn store_of(num_vars: usize, tables: &[Vec<Fr>]) -> VirtualPolynomialStore<Fr> {
    let mut store = VirtualPolynomialStore::new(num_vars);
    for t in tables {
        store.allocate_polynomial(t);
    }
    store
}

fn sumcheck_entry(num_vars: usize, tables: &[Vec<Fr>], expr: VirtualPolyExpr<Fr>, claimed: Fr, domain: &[u8]) -> Value {
    let mut store = store_of(num_vars, tables);
    let href = store.new_virtual_from_expr(expr);
    let mut t = Transcript::new(domain);
    let (proof, claim) = SumcheckProof::prove(num_vars, &store, &href, claimed, &mut t);
    json!({"claimed_sum": h(&claimed), "r_polys": polys(&proof.r_polys), "point": hs(&claim.point),
           "evaluation": h(&claim.evaluation), "state_end": hex(&t.state)})
}

fn zerocheck_entry(g2v: &[u64]) -> Value {
    // zerocheck.rs:85-211: g1 = i, g2 = i^2 (or a perturbed copy), h = g1 * g1 - g2
    let g1: Vec<Fr> = (0..8u64).map(Fr::from).collect();
    let g2: Vec<Fr> = g2v.iter().map(|v| Fr::from(*v)).collect();
    let mut store = store_of(3, &[g1, g2]);
    let e = VirtualPolyExpr::Input(0) * VirtualPolyExpr::Input(0) - VirtualPolyExpr::Input(1);
    let href = store.new_virtual_from_expr(e);
    let mut t = Transcript::new(b"zerocheck_test");
    // z is drawn inside prove; replay the draws on a copy of the transcript to report them
    let mut t2 = Transcript::new(b"zerocheck_test");
    let z: Vec<Fr> = (0..3).map(|_| t2.draw_field_element::<Fr>()).collect();
    let (proof, claim) = ZeroCheckProof::prove(&mut store, &href, &mut t);
    json!({"g2": g2v, "r_polys": polys(&proof.sumcheck_proof.r_polys), "point": hs(&claim.point),
           "evaluation": h(&claim.evaluation), "z": hs(&z), "state_end": hex(&t.state)})
}

fn fibonacci() -> (TransitionCircuit<Fr>, TraceWitness<Fr>) {
    // hyperplonk/tests/test_basic_proof.rs:17-52
    let mut circuit: TransitionCircuit<Fr> = TransitionCircuit::new(8);
    let s1 = circuit.allocate_state_cell();
    let s2 = circuit.allocate_state_cell();
    circuit.enforce_boundary_constraint(0, s1.current.to_expr());
    circuit.enforce_boundary_constraint(0, s2.current.to_expr() - VirtualPolyExpr::Const(Fr::from(1u64)));
    circuit.enforce_constraint(s2.next.to_expr() - (s1.current.to_expr() + s2.current.to_expr()));
    circuit.enforce_constraint(s1.next.to_expr() - s2.current.to_expr());
    let mut w: Vec<Vec<Fr>> = vec![vec![Fr::zero(); circuit.num_rows()]; circuit.num_cols()];
    for row in 0..circuit.num_rows() {
        if row == 0 {
            w[s1.current.col][0] = Fr::from(0u64);
            w[s2.current.col][0] = Fr::from(1u64);
            w[s1.next.col][0] = Fr::from(1u64);
            w[s2.next.col][0] = Fr::from(1u64);
        } else {
            w[s1.current.col][row] = w[s1.next.col][row - 1];
            w[s2.current.col][row] = w[s2.next.col][row - 1];
            w[s1.next.col][row] = w[s2.current.col][row];
            w[s2.next.col][row] = w[s2.current.col][row] + w[s1.current.col][row];
        }
    }
    (circuit, TraceWitness(w))
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let golden_path = args.get(1).cloned().unwrap_or_else(|| "tests/golden/golden.json".to_string());
    let out_path = args.get(2).cloned().unwrap_or_else(|| "tests/golden/reference.json".to_string());
    let golden: Value = serde_json::from_str(&std::fs::read_to_string(&golden_path).unwrap()).unwrap();
    let mut out = Map::new();
    out.insert("_generator".into(), json!("tools/refvec (the unmodified reference on arkworks 0.5.0 / blake3 1.8.2)"));

    out.insert("transcript".into(), transcript_vectors());

    // sumcheck.rs:159-230
    let g1: Vec<Fr> = (0..8u64).map(|i| Fr::from((i & 1) + 2 * ((i >> 1) & 1) + 3 * ((i >> 2) & 1))).collect();
    let g2: Vec<Fr> = (0..8u64).map(|i| Fr::from((i & 1) * 2 * ((i >> 1) & 1) + 3 * (i & 1) * ((i >> 2) & 1))).collect();
    let cs: Fr = g1.iter().zip(g2.iter()).map(|(a, b)| *a * *b).sum();
    out.insert(
        "sumcheck_test".into(),
        sumcheck_entry(3, &[g1, g2], VirtualPolyExpr::Input(0) * VirtualPolyExpr::Input(1), cs, b"sumcheck_test"),
    );

    out.insert("zerocheck_test".into(), zerocheck_entry(&[0, 1, 4, 9, 16, 25, 36, 49]));
    out.insert("zerocheck_test_not_zero".into(), zerocheck_entry(&[0, 1, 4, 9, 16, 25, 36, 50]));

    // seeded degree-3 product and mixed expression over 6 variables: tables from golden.json
    let tabs: Vec<Vec<Fr>> = golden["product3_n6"]["tables"].as_array().unwrap().iter().map(frs).collect();
    let cs3: Fr = (0..64).map(|i| tabs[0][i] * tabs[1][i] * tabs[2][i]).sum();
    let h3 = (VirtualPolyExpr::Input(0) * VirtualPolyExpr::Input(1)) * VirtualPolyExpr::Input(2);
    let mut e = sumcheck_entry(6, &tabs, h3, cs3, b"sumcheck_bench");
    e["tables"] = golden["product3_n6"]["tables"].clone();
    out.insert("product3_n6".into(), e);
    // make_golden.py: e_add(e_sub(e_mul(in0, in1), in3), e_mul(const 7, e_mul(in2, in2))), false claim 123
    let hm = (VirtualPolyExpr::Input(0) * VirtualPolyExpr::Input(1) - VirtualPolyExpr::Input(3))
        + VirtualPolyExpr::Const(Fr::from(7u64)) * (VirtualPolyExpr::Input(2) * VirtualPolyExpr::Input(2));
    let mut e = sumcheck_entry(6, &tabs, hm, Fr::from(123u64), b"mixed");
    e.as_object_mut().unwrap().remove("claimed_sum");
    out.insert("mixed_n6".into(), e);

    // eq_eval.rs:53-75
    let pt5 = frs(&golden["eq_n5"]["point"]);
    out.insert("eq_n5".into(), json!({"point": hs(&pt5), "table": hs(&fast_eq_eval_hypercube(5, &pt5))}));

    // kzg.rs:119-151 on the fixed SRS
    let kzg = fixed_kzg(4);
    let poly = vec![Fr::from(2u64), Fr::from(1u64), Fr::from(3u64)];
    let com = kzg.commit(&poly);
    let op = kzg.open(&poly, Fr::from(5u64));
    let tau = Fr::from(0x1234567890ABCDEF1234567890ABCDEFu128);
    out.insert(
        "kzg_test".into(),
        json!({"g": point(&kzg.g1), "tau": h(&tau), "srs": kzg.g1_points.iter().map(point).collect::<Vec<_>>(),
               "commitment": point(&com), "commitment_bytes": ser(&com), "y": h(&op.y), "proof": point(&op.proof),
               "identity_bytes": ser(&G1Projective::zero())}),
    );
    let kz64 = fixed_kzg(63);
    let sc = frs(&golden["msm64"]["scalars"]);
    out.insert("msm64".into(), json!({"scalars": hs(&sc), "result": point(&kz64.commit(&sc))}));

    // mlpcs.rs:220-243, ipa.rs:214-298
    let z = Fr::zero();
    let o = Fr::one();
    out.insert(
        "pr".into(),
        json!({"r000": hs(&MLEvalProof::<Bn254>::compute_pr(&[z, z, z]).coeffs),
               "r101": hs(&MLEvalProof::<Bn254>::compute_pr(&[o, z, o]).coeffs)}),
    );
    let f = |v: &[u64]| v.iter().map(|x| Fr::from(*x)).collect::<Vec<Fr>>();
    out.insert(
        "s_poly".into(),
        json!({"a123_b456": hs(&InnerProductProof::<Bn254>::compute_s_polynomial(&f(&[1, 2, 3]), &f(&[4, 5, 6])).coeffs),
               "a123_b45": hs(&InnerProductProof::<Bn254>::compute_s_polynomial(&f(&[1, 2, 3]), &f(&[4, 5])).coeffs)}),
    );

    // MLEvalProof::prove (mlpcs.rs:83-124) on golden.json's 5-variable inputs
    if let Some(m) = golden.get("mlpcs_n5") {
        let poly = frs(&m["poly"]);
        let pt = frs(&m["point"]);
        let kz = fixed_kzg(64);
        let mut t = Transcript::new(b"mlpcs_golden");
        let com = kz.commit(&poly);
        let pf = MLEvalProof::<Bn254>::prove(&poly, &pt, &kz, &mut t);
        let opening = |o: &quill_pcs::kzg::KZGOpeningProof<Bn254>| json!({"x": h(&o.x), "y": h(&o.y), "proof_bytes": ser(&o.proof)});
        out.insert(
            "mlpcs_n5".into(),
            json!({"poly": hs(&poly), "point": hs(&pt), "commitment_bytes": ser(&com), "evaluation": h(&pf.evaluation),
                   "s_comm_bytes": ser(&pf.s_comm), "poly_opening": opening(&pf.poly_opening),
                   "poly_opening_inv": opening(&pf.poly_opening_inv), "s_opening": opening(&pf.s_opening),
                   "s_opening_inv": opening(&pf.s_opening_inv), "state_end": hex(&t.state)}),
        );
    }

    // HyperPlonk::prove on the Fibonacci circuit (test_basic_proof.rs:137-164).  The reference keeps its transcript
    // private, so only proof fields are emitted (golden.json's state_end stays pinned through them: every later
    // challenge is a function of the fields below).
    let (circuit, witness) = fibonacci();
    let pcs = fixed_kzg(64);
    let hp = HyperPlonk::preprocess(vec![circuit.clone()], &pcs);
    let proof = hp.prove(&pcs, &vec![witness]);
    let tp = &proof.trace_proofs[0];
    out.insert(
        "hyperplonk".into(),
        json!({"fibonacci": {
            "witness_commitment": ser(&proof.witness_commitment[0]),
            "zc_round0": hs(&tp.zero_check_proof.sumcheck_proof.r_polys[0].coeffs),
            "perm_point": hs(&tp.opening_id.evaluation_point),
        }}),
    );

    let _ = (Fq::one(), G1Affine::identity());
    std::fs::write(&out_path, serde_json::to_string_pretty(&Value::Object(out)).unwrap()).unwrap();
    println!("wrote {}", out_path);
}
