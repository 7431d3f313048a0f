//! `degree_reduce` -- dist-primitive/src/degree_reduce.rs:29-41: gather -> unpack2 -> pack_from_public -> scatter.
//! (`degree_reduce_many`, :10-26, is used by `c_acc_product_and_share` only and runs inside scz_c_acc_product_and_share_dev.)
use crate::elements::{fr_from, SczFr};
use crate::net::GpuNet;
use mpc_net::{MPCNetError, MultiplexedStreamID};
use scz_sys::*;
use secret_sharing::pss::PackedSharingParams;

pub async fn degree_reduce<F: SczFr, Net: GpuNet>(
    shares: F,
    pp: &PackedSharingParams<F>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<F, MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let dpp = p.pp(pp.l)?;
    let d_in = p.upload(&[shares])?;
    let d_out = p.alloc(SCZ_FR_BYTES)?;
    crate::check(p, unsafe { scz_degree_reduce_dev(p.ctx(), dpp, d_in.ptr, d_out.ptr) })?;
    Ok(fr_from::<F>(&p.download::<u64>(&d_out, 4)?))
}
