//! `KZG::commit` / `KZG::open` (pcs/src/kzg.rs:61-96) on the device.
use crate::device::{fr_bytes, frs_bytes, pack_g1, unpack_fr, unpack_g1, Device};
use ark_bn254::{Bn254, Fr, G1Projective};
use ark_ec::CurveGroup;
use core::ffi::c_void;
use quill_b200_sys as sys;
use quill_pcs::kzg::{KZGOpeningProof, KZG};
use std::sync::Arc;

/// The reference's `KZG<Bn254>` plus its G1 powers resident in HBM.  `inner` keeps every pub field of the reference
/// struct (max_degree, g1, g2, g1_points, g2_points), so verification and anything else that reads the CRS is untouched.
pub struct KzgB200 {
    pub inner: KZG<Bn254>,
    pub device: Arc<Device>,
    pub(crate) srs: *mut sys::qz_srs,
}
unsafe impl Send for KzgB200 {}

impl KzgB200 {
    /// Upload the SRS once.  This is the affine normalisation the reference repeats on EVERY commit (kzg.rs:67-71).
    /// `precompute`: also store the window multiples 2^(c w) P_i (W x the SRS in HBM) for the shared-bucket MSM.
    pub fn new(inner: KZG<Bn254>, device: Arc<Device>, precompute: bool) -> Self {
        let affine = G1Projective::normalize_batch(&inner.g1_points);
        let mut xy = Vec::with_capacity(64 * affine.len());
        for p in &affine {
            xy.extend_from_slice(&pack_g1(p));
        }
        let mut srs = core::ptr::null_mut();
        device.check(unsafe { sys::qz_srs_upload(device.ctx, xy.as_ptr(), affine.len(), &mut srs) });
        if precompute {
            device.check(unsafe { sys::qz_srs_precompute(device.ctx, srs, 0) });
        }
        KzgB200 { inner, device, srs }
    }

    /// kzg.rs:61-73.  Panics with the reference's message when the polynomial is longer than the SRS.
    pub fn commit(&self, polynomial: &[Fr]) -> G1Projective {
        let mut out = [0u8; 64];
        self.device.check(unsafe {
            sys::qz_kzg_commit(self.device.ctx, self.srs, frs_bytes(polynomial) as *const c_void, polynomial.len(), 0, out.as_mut_ptr())
        });
        unpack_g1(&out)
    }

    /// The bare MSM seam, `E::G1::msm_unchecked(&affine, scalars)` (kzg.rs:72): zips to the shorter slice.
    pub fn msm_unchecked(&self, scalars: &[Fr]) -> G1Projective {
        let mut out = [0u8; 64];
        self.device.check(unsafe {
            sys::qz_msm(self.device.ctx, self.srs, frs_bytes(scalars) as *const c_void, scalars.len(), 0, out.as_mut_ptr())
        });
        unpack_g1(&out)
    }

    /// kzg.rs:75-96: y = p(x), proof = commit((p - y) / (X - x)).
    pub fn open(&self, polynomial: &[Fr], x: Fr) -> KZGOpeningProof<Bn254> {
        let (mut y, mut proof) = ([0u8; 32], [0u8; 64]);
        self.device.check(unsafe {
            sys::qz_kzg_open(self.device.ctx, self.srs, frs_bytes(polynomial) as *const c_void, polynomial.len(), 0,
                             fr_bytes(&x), y.as_mut_ptr(), proof.as_mut_ptr())
        });
        KZGOpeningProof { x, y: unpack_fr(&y), proof: unpack_g1(&proof) }
    }

    /// kzg.rs:98-108 is verifier-side and stays on the reference's code.
    pub fn verify(&self, commitment: &G1Projective, proof: &KZGOpeningProof<Bn254>) -> bool {
        self.inner.verify(commitment, proof)
    }
}
impl Drop for KzgB200 {
    fn drop(&mut self) {
        unsafe { sys::qz_srs_free(self.srs) }
    }
}
