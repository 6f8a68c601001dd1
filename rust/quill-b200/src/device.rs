//! Context handle, status handling and the byte-level packing of field elements and points.
use ark_bn254::{Fq, Fr, G1Affine, G1Projective};
use ark_ec::{AffineRepr, CurveGroup};
use ark_ff::{BigInt, Zero};
use core::ffi::CStr;
use quill_b200_sys as sys;

// ark-ff's Fp is `Fp(pub BigInt<N>, PhantomData)`: 4 x u64 Montgomery limbs, little-endian
const _: () = assert!(core::mem::size_of::<Fr>() == 32 && core::mem::size_of::<Fq>() == 32);

/// One GPU, one stream (`qz_ctx`).  One in-flight call per device, like the reference's `&mut Transcript` discipline.
pub struct Device {
    pub(crate) ctx: *mut sys::qz_ctx,
}
unsafe impl Send for Device {}

impl Device {
    pub fn new(device: i32) -> Self {
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe { sys::qz_ctx_create(device, core::ptr::null_mut(), &mut ctx) };
        if rc != sys::QZ_OK {
            let why = unsafe { CStr::from_ptr(sys::qz_status_str(rc)) }.to_string_lossy().into_owned();
            panic!("quill_b200: qz_ctx_create({device}) failed: {why}");
        }
        Device { ctx }
    }
    /// The reference's prover panics on bad input (kzg.rs:62-65, virtual_polynomial.rs:162-166, `.unwrap()` on
    /// inverses); the library returns a status instead, which is turned back into a panic here.
    pub(crate) fn check(&self, rc: i32) {
        if rc == sys::QZ_OK {
            return;
        }
        if rc == sys::QZ_ERR_DEGREE {
            panic!("Polynomial degree exceeds max degree"); // kzg.rs:62-65, verbatim
        }
        let status = unsafe { CStr::from_ptr(sys::qz_status_str(rc)) }.to_string_lossy().into_owned();
        let detail = unsafe { CStr::from_ptr(sys::qz_last_error(self.ctx)) }.to_string_lossy().into_owned();
        panic!("quill_b200: {status}: {detail}");
    }
    /// Join the box-wide communicator (one process per GPU): `id` from `unique_id()` on rank 0, broadcast by the launcher.
    pub fn comm_init(&self, id: &[u8; 128], rank: i32, nranks: i32) {
        self.check(unsafe { sys::qz_comm_init(self.ctx, id.as_ptr(), rank, nranks) });
    }
    pub fn unique_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        assert_eq!(unsafe { sys::qz_comm_unique_id(id.as_mut_ptr()) }, sys::QZ_OK, "NCCL unavailable");
        id
    }
}
impl Drop for Device {
    fn drop(&mut self) {
        unsafe { sys::qz_ctx_destroy(self.ctx) }
    }
}

pub(crate) fn fr_bytes(x: &Fr) -> *const u8 {
    x as *const Fr as *const u8
}
pub(crate) fn frs_bytes(xs: &[Fr]) -> *const u8 {
    xs.as_ptr() as *const u8
}
pub(crate) fn frs_bytes_mut(xs: &mut [Fr]) -> *mut u8 {
    xs.as_mut_ptr() as *mut u8
}
fn limbs(b: &[u8]) -> BigInt<4> {
    let mut l = [0u64; 4];
    for i in 0..4 {
        l[i] = u64::from_le_bytes(b[8 * i..8 * i + 8].try_into().unwrap());
    }
    BigInt::new(l)
}
/// `G1Affine { x, y, infinity }` is not `repr(C)`: pack x ‖ y Montgomery limbs, all-zero = the point at infinity
pub(crate) fn pack_g1(p: &G1Affine) -> [u8; 64] {
    let mut o = [0u8; 64];
    if let Some((x, y)) = p.xy() {
        for i in 0..4 {
            o[8 * i..8 * i + 8].copy_from_slice(&x.0 .0[i].to_le_bytes());
            o[32 + 8 * i..40 + 8 * i].copy_from_slice(&y.0 .0[i].to_le_bytes());
        }
    }
    o
}
pub(crate) fn pack_g1_projective(p: &G1Projective) -> [u8; 64] {
    pack_g1(&p.into_affine())
}
pub(crate) fn unpack_g1(b: &[u8]) -> G1Projective {
    if b[..64].iter().all(|v| *v == 0) {
        return G1Projective::zero();
    }
    // the library returns canonical Montgomery residues of a point on the curve; no subgroup check is needed (G1 of
    // BN254 has cofactor 1)
    G1Affine::new_unchecked(Fq::new_unchecked(limbs(&b[..32])), Fq::new_unchecked(limbs(&b[32..64]))).into()
}
pub(crate) fn unpack_fr(b: &[u8]) -> Fr {
    Fr::new_unchecked(limbs(&b[..32]))
}
