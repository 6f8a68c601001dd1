//! `ZeroCheckProof::prove` (hyperplonk/src/piops/zerocheck.rs:14-49).
use crate::device::{frs_bytes, frs_bytes_mut, Device};
use crate::expr::flatten;
use crate::sumcheck::{round_polys, table_ptrs, MC};
use ark_bn254::Fr;
use ark_std::Zero;
use quill_b200_sys as sys;
use quill_hyperplonk::piops::sumcheck::SumcheckProof;
use quill_hyperplonk::piops::zerocheck::ZeroCheckProof;
use quill_hyperplonk::utils::eq_eval::fast_eq_eval_hypercube;
use quill_hyperplonk::utils::virtual_polynomial::{VirtualPolynomialRef, VirtualPolynomialStore};
use quill_pcs::EvaluationClaim;
use quill_transcript::transcript::Transcript;

/// The library draws z, builds eq(., z) (or never materialises it: eq-factored rounds), proves sum h * eq = 0 and
/// divides the final claim by eq(z, r).  The reference also leaves two things behind in the caller's store -- the eq
/// table as a new polynomial and h_hat = h * eq as a new virtual polynomial (zerocheck.rs:27-29) -- and later code may
/// index past them (`proof.rs:184` opens a fresh store, but callers are free not to), so both are appended here too;
/// pass `mirror_store_side_effects = false` to skip the 2^n-element eq table when nothing reads it.
pub fn prove(
    dev: &Device,
    store: &mut VirtualPolynomialStore<Fr>,
    h: &VirtualPolynomialRef,
    transcript: &mut Transcript,
    mirror_store_side_effects: bool,
) -> (ZeroCheckProof<Fr>, EvaluationClaim<Fr>) {
    assert_eq!(transcript.state.len(), 32);
    let num_vars = store.num_vars();
    let (nodes, consts) = flatten(&store.virtual_polys[h.index]);
    let tabs = table_ptrs(store);
    let mut coeffs = vec![Fr::zero(); num_vars.max(1) * MC];
    let mut lens = vec![0u32; num_vars.max(1)];
    let mut point = vec![Fr::zero(); num_vars.max(1)];
    let mut z = vec![Fr::zero(); num_vars.max(1)];
    let mut eval = Fr::zero();
    dev.check(unsafe {
        sys::qz_zerocheck_prove(dev.ctx, num_vars, tabs.len(), tabs.as_ptr(), 0, nodes.as_ptr(), nodes.len(), frs_bytes(&consts),
                                consts.len(), transcript.state.as_mut_ptr(), MC, frs_bytes_mut(&mut coeffs), lens.as_mut_ptr(),
                                frs_bytes_mut(&mut point), &mut eval as *mut Fr as *mut u8, frs_bytes_mut(&mut z))
    });
    point.truncate(num_vars);
    z.truncate(num_vars);
    if mirror_store_side_effects {
        let eq_evals = fast_eq_eval_hypercube(num_vars, z.as_slice());
        let eq_poly_index = store.allocate_polynomial(&eq_evals);
        let h_hat = store.new_virtual_from_virtual(h);
        store.mul_in_place(&h_hat, &eq_poly_index);
    }
    let sumcheck_proof = SumcheckProof { num_vars, claimed_sum: Fr::zero(), r_polys: round_polys(&coeffs, &lens[..num_vars]) };
    (ZeroCheckProof { num_vars, sumcheck_proof }, EvaluationClaim { point, evaluation: eval })
}
