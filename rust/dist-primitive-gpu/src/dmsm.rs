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
