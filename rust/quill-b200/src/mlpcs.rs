//! `impl MultilinearPCS<Fr>` (pcs/src/lib.rs:26-41) with the associated types of the reference's impl for `KZG<E>`
//! (pcs/src/mlpcs.rs:174-207): `HyperPlonk<Fr, C, KzgB200>` and every PIOP generic over `PCS` accept it unchanged.
use crate::device::{frs_bytes, unpack_fr, unpack_g1, Device};
use crate::kzg::KzgB200;
use ark_bn254::{Bn254, Fr, G1Projective};
use core::ffi::c_void;
use quill_b200_sys as sys;
use quill_pcs::kzg::{KZGOpeningProof, KZG};
use quill_pcs::mlpcs::MLEvalProof;
use quill_pcs::MultilinearPCS;
use quill_transcript::transcript::Transcript;
use std::sync::Arc;

/// `MLEvalProof::prove` (mlpcs.rs:83-124) as ONE device call: P_r (the eq table of the point), the evaluation, the S
/// polynomial (NTT), commit(S), the transcript schedule of mlpcs.rs:100-105 and the four KZG openings -- five MSMs with
/// every polynomial resident in HBM.  The 32 + 64 + 4 x 128 output bytes are unpacked into the reference's struct.
pub fn ml_eval_prove(kzg: &KzgB200, poly: &[Fr], eval_point: &[Fr], transcript: &mut Transcript) -> MLEvalProof<Bn254> {
    assert_eq!(transcript.state.len(), 32, "Transcript.state is a 32-byte blake3 digest");
    let (mut evaluation, mut s_comm, mut openings) = ([0u8; 32], [0u8; 64], [0u8; 512]);
    kzg.device.check(unsafe {
        sys::qz_mlpcs_open(kzg.device.ctx, kzg.srs, frs_bytes(poly) as *const c_void, poly.len(), 0, frs_bytes(eval_point),
                           eval_point.len(), transcript.state.as_mut_ptr(), evaluation.as_mut_ptr(), s_comm.as_mut_ptr(),
                           openings.as_mut_ptr())
    });
    // 4 x [x (32) ‖ y (32) ‖ proof (64)]: poly_opening, poly_opening_inv, s_opening, s_opening_inv (mlpcs.rs:109-113)
    let opening = |i: usize| -> KZGOpeningProof<Bn254> {
        let o = &openings[128 * i..128 * (i + 1)];
        KZGOpeningProof { x: unpack_fr(&o[..32]), y: unpack_fr(&o[32..64]), proof: unpack_g1(&o[64..128]) }
    };
    MLEvalProof {
        evaluation_point: eval_point.to_vec(),
        evaluation: unpack_fr(&evaluation),
        s_comm: unpack_g1(&s_comm),
        poly_opening: opening(0),
        poly_opening_inv: opening(1),
        s_opening: opening(2),
        s_opening_inv: opening(3),
    }
}

impl MultilinearPCS<Fr> for KzgB200 {
    type CRS = KzgB200;
    type Commitment = G1Projective;
    type Proof = MLEvalProof<Bn254>;

    fn trusted_setup(degree: usize) -> Self::CRS {
        // mlpcs.rs:179-182 draws the CRS from thread_rng(); the G1 powers are then uploaded once
        let mut rng = rand::thread_rng();
        KzgB200::new(KZG::<Bn254>::trusted_setup(degree, &mut rng), Arc::new(Device::new(0)), true)
    }
    fn max_degree(&self) -> usize {
        self.inner.max_degree
    }
    fn commit(&self, poly: &[Fr]) -> Self::Commitment {
        KzgB200::commit(self, poly)
    }
    fn open(&self, poly: &[Fr], eval_point: &[Fr], transcript: &mut Transcript) -> Self::Proof {
        ml_eval_prove(self, poly, eval_point, transcript)
    }
    fn verify(&self, commitment: &Self::Commitment, proof: &Self::Proof, transcript: &mut Transcript) -> bool {
        proof.verify(commitment, &self.inner, transcript) // verifier side: the reference's own code (mlpcs.rs:126-161)
    }
}
