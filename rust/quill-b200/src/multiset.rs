//! `MultisetEqualityProof::prove` (hyperplonk/src/piops/multiset_check.rs:28-182) over the device primitives: the two
//! logup denominator tables come from `qz_logup_denominators` (one batched inversion instead of 2^n `inverse().unwrap()`
//! calls, :43-95), the commitments and openings from the PCS (`KzgB200`), the batched sumcheck from `sumcheck::prove`.
//! The transcript schedule is the reference's, line for line.
use crate::device::{fr_bytes, frs_bytes, frs_bytes_mut, Device};
use crate::expr::{flatten, flatten_pair};
use crate::kzg::KzgB200;
use crate::sumcheck;
use ark_bn254::Fr;
use ark_std::{One, Zero};
use core::ffi::c_void;
use quill_b200_sys as sys;
use quill_hyperplonk::piops::multiset_check::{LookupMode, MultisetEqualityProof};
use quill_hyperplonk::utils::eq_eval::fast_eq_eval_hypercube;
use quill_hyperplonk::utils::virtual_polynomial::{VirtualPolyExpr, VirtualPolynomialRef, VirtualPolynomialStore};
use quill_pcs::MultilinearPCS;
use quill_transcript::transcript::Transcript;

/// out[i] = m(row_i) / (gamma + h(row_i)) over the store's tables (multiset_check.rs:43-95).  Panics where the reference's
/// `.inverse().unwrap()` would (a zero denominator).
pub fn logup_denominators(
    dev: &Device,
    store: &VirtualPolynomialStore<Fr>,
    h: &VirtualPolyExpr<Fr>,
    m: Option<&VirtualPolyExpr<Fr>>,
    gamma: Fr,
) -> Vec<Fr> {
    let tabs = sumcheck::table_ptrs(store);
    let mut out = vec![Fr::zero(); 1 << store.num_vars];
    let (nh, nm, consts) = match m {
        Some(m) => flatten_pair(h, m),
        None => {
            let (nh, c) = flatten(h);
            (nh, Vec::new(), c)
        }
    };
    dev.check(unsafe {
        sys::qz_logup_denominators(dev.ctx, store.num_vars, tabs.len(), tabs.as_ptr(), 0, nh.as_ptr(), nh.len(),
                                   if nm.is_empty() { core::ptr::null() } else { nm.as_ptr() }, nm.len(), frs_bytes(&consts),
                                   consts.len(), fr_bytes(&gamma), frs_bytes_mut(&mut out) as *mut c_void, 0)
    });
    out
}

pub fn prove(
    dev: &Device,
    store: &mut VirtualPolynomialStore<Fr>,
    h_left: &VirtualPolynomialRef,
    h_right: &VirtualPolynomialRef,
    transcript: &mut Transcript,
    pcs: &KzgB200,
    mode: LookupMode,
    multiplicities: Option<&VirtualPolynomialRef>,
) -> (MultisetEqualityProof<Fr, KzgB200>, Vec<Fr>) {
    let num_vars = store.num_vars();
    let logup_eval_point = transcript.draw_field_element::<Fr>(); // :40

    let m_expr = match mode {
        LookupMode::Subset => {
            assert!(multiplicities.is_some(), "Multiplicities polynomial must be provided in subset mode");
            Some(store.virtual_polys[multiplicities.unwrap().index].clone())
        }
        LookupMode::Equality => {
            assert!(multiplicities.is_none(), "Multiplicities polynomial must not be provided in equality mode");
            None
        }
    };
    let h_left_expr = store.virtual_polys[h_left.index].clone();
    let h_right_expr = store.virtual_polys[h_right.index].clone();
    let log_derivative_left_evals = logup_denominators(dev, store, &h_left_expr, None, logup_eval_point); // :43-53
    let log_derivative_right_evals = logup_denominators(dev, store, &h_right_expr, m_expr.as_ref(), logup_eval_point); // :55-95

    let commitment_left = MultilinearPCS::commit(pcs, &log_derivative_left_evals); // :98-101
    let commitment_right = MultilinearPCS::commit(pcs, &log_derivative_right_evals);
    transcript.append_serializable(&commitment_left);
    transcript.append_serializable(&commitment_right);

    let lambda = transcript.draw_field_element::<Fr>(); // :104-105
    let alpha = transcript.draw_field_element::<Fr>();

    let denom_left_index = store.allocate_polynomial(&log_derivative_left_evals); // :108-109
    let denom_right_index = store.allocate_polynomial(&log_derivative_right_evals);

    let m = m_expr.unwrap_or(VirtualPolyExpr::Const(Fr::one())); // :128-131
    let zerocheck_expr = denom_left_index.to_expr::<Fr>() * (VirtualPolyExpr::Const(logup_eval_point) + h_left_expr)
        - VirtualPolyExpr::Const(Fr::one())
        + VirtualPolyExpr::Const(lambda)
            * (denom_right_index.to_expr::<Fr>() * (VirtualPolyExpr::Const(logup_eval_point) + h_right_expr) - m); // :132-141

    let zerocheck_random_point = (0..num_vars).map(|_| transcript.draw_field_element::<Fr>()).collect::<Vec<Fr>>(); // :144-146
    // the eq table as a store polynomial, on the device (qz_eq_table) rather than by fast_eq_eval_hypercube on the host
    let mut eq_evals = vec![Fr::zero(); 1 << num_vars];
    dev.check(unsafe {
        sys::qz_eq_table(dev.ctx, num_vars, frs_bytes(&zerocheck_random_point), frs_bytes_mut(&mut eq_evals) as *mut c_void, 0)
    });
    debug_assert_eq!(eq_evals, fast_eq_eval_hypercube(num_vars, zerocheck_random_point.as_slice()));
    let eq_poly_index = store.allocate_polynomial(&eq_evals); // :149-150
    let h_hat = store.new_virtual_from_expr(zerocheck_expr); // :152-153
    store.mul_in_place(&h_hat, &eq_poly_index);
    store.mul_const_in_place(&h_hat, alpha); // :156-158
    store.add_in_place(&h_hat, &denom_left_index);
    store.sub_in_place(&h_hat, &denom_right_index);

    let (sumcheck_proof, sumcheck_evaluation_claim) = sumcheck::prove(dev, num_vars, store, &h_hat, Fr::zero(), transcript); // :162-163
    let evaluation_point = sumcheck_evaluation_claim.point;
    let opening_proof_denom_left = MultilinearPCS::open(pcs, &log_derivative_left_evals, &evaluation_point, transcript); // :167-170
    let opening_proof_denom_right = MultilinearPCS::open(pcs, &log_derivative_right_evals, &evaluation_point, transcript);
    (
        MultisetEqualityProof {
            denom_left_commitment: commitment_left,
            denom_right_commitment: commitment_right,
            sumcheck_proof,
            opening_proof_denom_left,
            opening_proof_denom_right,
        },
        evaluation_point,
    )
}
