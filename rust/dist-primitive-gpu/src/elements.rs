//! Element layouts at the C boundary (include/scz.h header comment).
//!   Fr          = `Fp<MontBackend<FrConfig, 4>, 4>(BigInt([u64; 4]))`: Montgomery limbs, passed as they lie in memory
//!   G1Affine    = `{x, y, infinity}` is repr(Rust): repacked once into x | y (12 x u64) + an `infinity` byte mask
//!   G1 (result) = Jacobian X | Y | Z, 18 x u64 = `Projective {x, y, z}`
use ark_bls12_381::{Fq, Fr, G1Affine, G1Projective};
use ark_ec::CurveGroup;
use ark_ff::{BigInt, FftField, Fp, PrimeField};
use core::ffi::c_void;

mod sealed {
    pub trait Sealed {}
    impl Sealed for ark_bls12_381::Fr {}
    impl Sealed for ark_bls12_381::G1Projective {}
}

/// The one scalar field libscz implements.  `F: FftField + SczFr` keeps the reference's generic signatures.
pub trait SczFr: FftField + sealed::Sealed {
    /// `&[F]` as the raw Montgomery limbs libscz reads (zero-copy)
    fn as_raw(v: &[Self]) -> *const c_void;
    fn as_raw_mut(v: &mut [Self]) -> *mut c_void;
}
impl SczFr for Fr {
    fn as_raw(v: &[Fr]) -> *const c_void {
        // Fp<MontBackend<_, 4>, 4> is a transparent wrapper chain around [u64; 4] (ark-ff 0.4.2 fields/models/fp/mod.rs)
        const _: () = assert!(core::mem::size_of::<Fr>() == scz_sys::SCZ_FR_BYTES);
        v.as_ptr() as *const c_void
    }
    fn as_raw_mut(v: &mut [Fr]) -> *mut c_void {
        v.as_mut_ptr() as *mut c_void
    }
}

/// The one group libscz implements (`d_msm` is only ever instantiated with G1: dpoly_comm.rs:265, examples/msm.rs:66,89).
pub trait SczG1: CurveGroup<ScalarField = Fr, Affine = G1Affine> + sealed::Sealed {
    fn from_jacobian_limbs(l: &[u64; 18]) -> Self;
    fn to_jacobian_limbs(&self) -> [u64; 18];
}
fn fq_from_limbs(l: &[u64]) -> Fq {
    // the limbs ARE the Montgomery representation: build the element without a conversion
    Fp::new_unchecked(BigInt::new([l[0], l[1], l[2], l[3], l[4], l[5]]))
}
fn fq_limbs(x: &Fq) -> [u64; 6] {
    (x.0).0
}
impl SczG1 for G1Projective {
    fn from_jacobian_limbs(l: &[u64; 18]) -> Self {
        G1Projective::new_unchecked(fq_from_limbs(&l[0..6]), fq_from_limbs(&l[6..12]), fq_from_limbs(&l[12..18]))
    }
    fn to_jacobian_limbs(&self) -> [u64; 18] {
        let mut o = [0u64; 18];
        o[0..6].copy_from_slice(&fq_limbs(&self.x));
        o[6..12].copy_from_slice(&fq_limbs(&self.y));
        o[12..18].copy_from_slice(&fq_limbs(&self.z));
        o
    }
}

/// x | y of every base plus the `infinity` flags, ready for `scz_msm_g1` / `scz_d_msm` / `scz_srs_from_host_levels`
pub struct PackedBases {
    pub xy: Vec<[u64; 12]>,
    pub infinity: Vec<u8>,
}
pub fn pack_affine(bases: &[G1Affine]) -> PackedBases {
    let mut xy = Vec::with_capacity(bases.len());
    let mut infinity = Vec::with_capacity(bases.len());
    for b in bases {
        let mut l = [0u64; 12];
        if !b.infinity {
            l[0..6].copy_from_slice(&fq_limbs(&b.x));
            l[6..12].copy_from_slice(&fq_limbs(&b.y));
        }
        xy.push(l);
        infinity.push(b.infinity as u8);
    }
    PackedBases { xy, infinity }
}

/// (Fr, Fr, Fr) round messages come back as 12 limbs each
pub fn triples_from_limbs<F: SczFr>(raw: &[[u64; 12]]) -> Vec<(F, F, F)> {
    raw.iter().map(|t| (fr_from::<F>(&t[0..4]), fr_from::<F>(&t[4..8]), fr_from::<F>(&t[8..12]))).collect()
}
pub fn pairs_from_limbs<F: SczFr>(raw: &[[u64; 8]]) -> Vec<(F, F)> {
    raw.iter().map(|t| (fr_from::<F>(&t[0..4]), fr_from::<F>(&t[4..8]))).collect()
}
pub fn fr_from<F: SczFr>(l: &[u64]) -> F {
    let mut out = [F::zero()];
    unsafe { core::ptr::copy_nonoverlapping(l.as_ptr() as *const u8, F::as_raw_mut(&mut out) as *mut u8, 32) };
    out[0]
}
#[allow(dead_code)]
fn _assert_prime_field<F: PrimeField>() {}
