//! `d_msm` -- same signature as dist-primitive/src/dmsm.rs:9-15.
use crate::elements::{pack_affine, SczG1};
use crate::net::GpuNet;
use core::ffi::c_void;
use mpc_net::{MPCNetError, MultiplexedStreamID};
use scz_sys::*;
use secret_sharing::pss::PackedSharingParams;

pub async fn d_msm<G: SczG1, Net: GpuNet>(
    bases: &Vec<Vec<G::Affine>>,
    scalars: &Vec<Vec<G::ScalarField>>,
    pp: &PackedSharingParams<G::ScalarField>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<Vec<G>, MPCNetError> {
    assert_eq!(bases.len(), scalars.len()); // dmsm.rs:16
    let p = net.gpu();
    let _g = p.lock();
    let dpp = p.pp(pp.l)?;
    // G1Affine is repr(Rust): repack x | y and fold `infinity` into x = y = 0 (what scz_d_msm's host path expects)
    let packed: Vec<_> = bases.iter().map(|b| pack_affine(b)).collect();
    let bp: Vec<*const c_void> = packed.iter().map(|b| b.xy.as_ptr() as *const c_void).collect();
    let bl: Vec<usize> = packed.iter().map(|b| b.xy.len()).collect();
    let sp: Vec<*const c_void> = scalars.iter().map(|s| s.as_ptr() as *const c_void).collect();
    let sl: Vec<usize> = scalars.iter().map(|s| s.len()).collect();
    let mut out = vec![[0u64; 18]; bases.len()];
    // local G::msm per batch entry (:19-24) -> gather -> leader: unpack2, sum of the l secrets, pack_from_public (:31-38)
    // -> scatter; a base / scalar length mismatch panics like the reference's unwrap() (:23), see crate::check
    crate::check(p, unsafe {
        scz_d_msm(p.ctx(), dpp, bp.as_ptr(), bl.as_ptr(), sp.as_ptr(), sl.as_ptr(), bases.len(), out.as_mut_ptr() as *mut c_void)
    })?;
    Ok(out.iter().map(G::from_jacobian_limbs).collect())
}

/// The L0 seam: `G::msm(bases, scalars)` (ark-ec VariableBaseMSM; call sites dmsm.rs:23, dpoly_comm.rs:242,274,457).
/// Returns `Err(min len)` on a length mismatch like ark-ec.
pub fn msm<G: SczG1, Net: GpuNet>(net: &Net, bases: &[G::Affine], scalars: &[G::ScalarField]) -> Result<G, usize> {
    if bases.len() != scalars.len() {
        return Err(bases.len().min(scalars.len()));
    }
    let p = net.gpu();
    let _g = p.lock();
    let b = pack_affine(bases);
    let mut out = [0u64; 18];
    let rc = unsafe {
        scz_msm_g1(p.ctx(), b.xy.as_ptr() as *const c_void, b.infinity.as_ptr(), bases.len(), scalars.as_ptr() as *const c_void,
                   scalars.len(), out.as_mut_ptr() as *mut c_void)
    };
    assert_eq!(rc, SCZ_OK, "scz_msm_g1: {}", p.last_error());
    Ok(G::from_jacobian_limbs(&out))
}

/// `d_msm` instantiated with G2 (`G: CurveGroup` in the reference's signature; no caller of the reference does it).  Bases are
/// repacked into x | y with Fq2 = c0 | c1 (24 x u64) + the infinity mask; results come back as Jacobian [u64; 36].
pub async fn d_msm_g2<Net: GpuNet>(
    bases: &Vec<Vec<ark_bls12_381::G2Affine>>,
    scalars: &Vec<Vec<ark_bls12_381::Fr>>,
    pp: &PackedSharingParams<ark_bls12_381::Fr>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<Vec<ark_bls12_381::G2Projective>, MPCNetError> {
    use ark_bls12_381::{Fq, Fq2, G2Projective};
    use ark_ff::{BigInt, Fp};
    assert_eq!(bases.len(), scalars.len()); // dmsm.rs:16
    let p = net.gpu();
    let _g = p.lock();
    let dpp = p.pp(pp.l)?;
    let limbs = |x: &Fq| (x.0).0;
    let fq = |l: &[u64]| -> Fq { Fp::new_unchecked(BigInt::new([l[0], l[1], l[2], l[3], l[4], l[5]])) };
    let fq2 = |l: &[u64]| Fq2::new(fq(&l[0..6]), fq(&l[6..12]));
    let mut d_b = Vec::new();
    let mut d_s = Vec::new();
    for (b, s) in bases.iter().zip(scalars.iter()) {
        if b.len() != s.len() {
            panic!("called `Result::unwrap()` on an `Err` value: {}", b.len().min(s.len())); // G::msm(..).unwrap(), dmsm.rs:23
        }
        let mut xy = vec![[0u64; 24]; b.len()];
        for (o, pt) in xy.iter_mut().zip(b.iter()) {
            if !pt.infinity {
                o[0..6].copy_from_slice(&limbs(&pt.x.c0));
                o[6..12].copy_from_slice(&limbs(&pt.x.c1));
                o[12..18].copy_from_slice(&limbs(&pt.y.c0));
                o[18..24].copy_from_slice(&limbs(&pt.y.c1));
            }
        }
        d_b.push(p.upload(&xy)?);
        d_s.push(p.upload(s)?);
    }
    let bp: Vec<*const c_void> = d_b.iter().map(|b| b.ptr as *const c_void).collect();
    let sp: Vec<*const c_void> = d_s.iter().map(|s| s.ptr as *const c_void).collect();
    let lens: Vec<usize> = bases.iter().map(|b| b.len()).collect();
    let d_o = p.alloc(bases.len().max(1) * SCZ_G2_JAC_BYTES)?;
    crate::check(p, unsafe { scz_d_msm_g2_dev(p.ctx(), dpp, bp.as_ptr(), sp.as_ptr(), lens.as_ptr(), bases.len(), d_o.ptr) })?;
    let raw = p.download::<u64>(&d_o, bases.len() * 36)?; // flat: `[u64; 36]` has no `Default` (arrays stop at 32)
    Ok(raw.chunks_exact(36).map(|j| G2Projective::new_unchecked(fq2(&j[0..12]), fq2(&j[12..24]), fq2(&j[24..36]))).collect())
}
