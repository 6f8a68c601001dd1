//! `quill-b200`: the reference's proving hot path on a B200, behind the reference's own Rust API.
//!
//! | reference item | here |
//! |---|---|
//! | `KZG::<Bn254>::commit` / `::open` (pcs/src/kzg.rs:61-96) | [`kzg::KzgB200::commit`], [`kzg::KzgB200::open`] |
//! | `impl MultilinearPCS<Fr> for KZG<Bn254>` (pcs/src/mlpcs.rs:174-207) | `impl MultilinearPCS<Fr> for KzgB200` ([`mlpcs`]) |
//! | `SumcheckProof::prove` (hyperplonk/src/piops/sumcheck.rs:28-114) | [`sumcheck::prove`] |
//! | `ZeroCheckProof::prove` (zerocheck.rs:14-49) | [`zerocheck::prove`] |
//! | `MultisetEqualityProof::prove` (multiset_check.rs:28-182) | [`multiset::prove`] |
//!
//! `HyperPlonk<Fr, C, KzgB200>` (proof/proof.rs:12) compiles unchanged against the trait impl; the three PIOP provers are
//! free functions with the reference's argument lists plus a leading `&Device`, to be called where the reference calls
//! `SumcheckProof::prove` (zerocheck.rs:31, multiset_check.rs:162), `ZeroCheckProof::prove` (proof.rs:179) and
//! `MultisetEqualityProof::prove` (permutation_check.rs:42).
//!
//! `Fr` crosses the ABI zero-copy: ark-ff stores an element as 4 x u64 Montgomery limbs, which is the library's layout.
pub mod device;
pub mod expr;
pub mod kzg;
pub mod mlpcs;
pub mod multiset;
pub mod sumcheck;
pub mod zerocheck;

pub use device::Device;
pub use kzg::KzgB200;
