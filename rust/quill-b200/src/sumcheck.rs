//! `SumcheckProof::prove` (hyperplonk/src/piops/sumcheck.rs:28-114): same arguments, return value and transcript effect.
use crate::device::{fr_bytes, frs_bytes, frs_bytes_mut, Device};
use crate::expr::flatten;
use ark_bn254::Fr;
use ark_poly::univariate::DensePolynomial;
use ark_poly::DenseUVPolynomial;
use ark_std::Zero;
use core::ffi::c_void;
use quill_b200_sys as sys;
use quill_hyperplonk::piops::sumcheck::SumcheckProof;
use quill_hyperplonk::utils::virtual_polynomial::{VirtualPolynomialRef, VirtualPolynomialStore};
use quill_pcs::EvaluationClaim;
use quill_transcript::transcript::Transcript;

pub(crate) const MC: usize = sys::QZ_MAX_ROUND_COEFFS;

/// The store's tables as host pointers (`DenseMultilinearExtension.evaluations` is a `Vec<Fr>`: zero-copy)
pub(crate) fn table_ptrs(store: &VirtualPolynomialStore<Fr>) -> Vec<*const c_void> {
    for p in &store.polynomials {
        assert_eq!(p.evaluations.len(), 1 << store.num_vars, "Input polynomial evaluations length does not match number of variables");
    }
    store.polynomials.iter().map(|p| p.evaluations.as_ptr() as *const c_void).collect()
}

/// Rows of `max_coeffs` zero-padded coefficients + lengths -> `Vec<DensePolynomial>` (already trimmed by the library
/// exactly as `DensePolynomial::from_coefficients_vec` would)
pub(crate) fn round_polys(coeffs: &[Fr], lens: &[u32]) -> Vec<DensePolynomial<Fr>> {
    lens.iter().enumerate().map(|(j, l)| DensePolynomial::from_coefficients_slice(&coeffs[j * MC..j * MC + *l as usize])).collect()
}

pub fn prove(
    dev: &Device,
    num_vars: usize,
    store: &VirtualPolynomialStore<Fr>,
    h: &VirtualPolynomialRef,
    claimed_sum: Fr,
    transcript: &mut Transcript,
) -> (SumcheckProof<Fr>, EvaluationClaim<Fr>) {
    assert_eq!(transcript.state.len(), 32);
    let (nodes, consts) = flatten(&store.virtual_polys[h.index]);
    let tabs = table_ptrs(store);
    let mut coeffs = vec![Fr::zero(); num_vars.max(1) * MC];
    let mut lens = vec![0u32; num_vars.max(1)];
    let mut point = vec![Fr::zero(); num_vars.max(1)];
    let mut eval = Fr::zero();
    dev.check(unsafe {
        sys::qz_sumcheck_prove(dev.ctx, num_vars, tabs.len(), tabs.as_ptr(), 0, nodes.as_ptr(), nodes.len(), frs_bytes(&consts),
                               consts.len(), fr_bytes(&claimed_sum), transcript.state.as_mut_ptr(), MC,
                               frs_bytes_mut(&mut coeffs), lens.as_mut_ptr(), frs_bytes_mut(&mut point),
                               &mut eval as *mut Fr as *mut u8)
    });
    point.truncate(num_vars);
    let r_polys = round_polys(&coeffs, &lens[..num_vars]);
    (SumcheckProof { num_vars, claimed_sum, r_polys }, EvaluationClaim { point, evaluation: eval })
}
