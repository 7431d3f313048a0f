//! `pss2ss` -- dist-primitive/src/unpack.rs:72-97: gather Fr -> leader unpack -> pack_single each secret -> scatter.
use crate::elements::{fr_from, SczFr};
use crate::net::GpuNet;
use mpc_net::{MPCNetError, MultiplexedStreamID};
use scz_sys::*;
use secret_sharing::pss::PackedSharingParams;

pub async fn pss2ss<F: SczFr, Net: GpuNet>(
    share: F,
    pp: &PackedSharingParams<F>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<Vec<F>, MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let dpp = p.pp(pp.l)?;
    let d_in = p.upload(&[share])?;
    let d_out = p.alloc(pp.l * SCZ_FR_BYTES)?;
    crate::check(p, unsafe { scz_pss2ss_dev(p.ctx(), dpp, d_in.ptr, d_out.ptr) })?;
    let raw = p.download::<u64>(&d_out, pp.l * 4)?;
    Ok(raw.chunks_exact(4).map(fr_from::<F>).collect())
}
