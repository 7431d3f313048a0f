//! Product accumulation ("prodcheck") -- dist-primitive/src/dacc_product.rs: acc_product :30-57, d_acc_product
//! :365-414, c_acc_product_and_share :66-292.
use crate::elements::{fr_from, SczFr};
use crate::net::GpuNet;
use mpc_net::{MPCNetError, MultiplexedStreamID};
use scz_sys::*;
use secret_sharing::pss::PackedSharingParams;

fn frs<F: SczFr>(raw: &[u64]) -> Vec<F> {
    raw.chunks_exact(4).map(fr_from::<F>).collect()
}

/// (v(x,0), v(x,1), v(1,x)) of the product tree (:41-56)
pub fn acc_product<F: SczFr, Net: GpuNet>(net: &Net, x: &Vec<F>) -> (Vec<F>, Vec<F>, Vec<F>) {
    let p = net.gpu();
    let _g = p.lock();
    let m = x.len();
    let d_x = p.upload(x).expect("upload");
    let d_t = p.alloc(2 * m * SCZ_FR_BYTES).expect("alloc");
    let rc = unsafe { scz_acc_product_dev(p.ctx(), d_x.ptr, m, d_t.ptr) };
    assert_eq!(rc, SCZ_OK, "scz_acc_product_dev: {}", p.last_error());
    let tree: Vec<F> = frs(&p.download::<u64>(&d_t, 2 * m * 4).expect("download"));
    // tree = x | products level by level | 0 (:34-39); the three outputs are its even / odd / upper-half entries (:41-56)
    let v0 = (0..m).map(|i| tree[2 * i]).collect();
    let v1 = (0..m).map(|i| tree[2 * i + 1]).collect();
    let v1x = tree[m..].to_vec();
    (v0, v1, v1x)
}

pub async fn d_acc_product<F: SczFr, Net: GpuNet>(
    inputs: &Vec<F>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<(Vec<F>, Option<Vec<F>>), MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let (m, n_parties) = (inputs.len(), net.n_parties());
    let d_x = p.upload(inputs)?;
    let d_sub = p.alloc(2 * m * SCZ_FR_BYTES)?;
    let d_top = p.alloc(2 * n_parties * SCZ_FR_BYTES)?;
    // subtree[2m - 1] is forced to zero BEFORE it is sent (:381,390): the leader's tree is built from what it receives
    crate::check(p, unsafe { scz_d_acc_product_dev(p.ctx(), d_x.ptr, m, d_sub.ptr, d_top.ptr) })?;
    let sub = frs(&p.download::<u64>(&d_sub, 2 * m * 4)?);
    let top = if net.is_leader() { Some(frs(&p.download::<u64>(&d_top, 2 * n_parties * 4)?)) } else { None };
    Ok((sub, top))
}

pub async fn c_acc_product_and_share<F: SczFr, Net: GpuNet>(
    shares: &Vec<F>,
    masks: &Vec<F>,
    unmask0: &Vec<F>,
    unmask1: &Vec<F>,
    unmask2: &Vec<F>,
    pp: &PackedSharingParams<F>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<(Vec<F>, Vec<F>, Vec<F>), MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let dpp = p.pp(pp.l)?;
    let len = shares.len();
    let (d_s, d_m, d_u0, d_u1, d_u2) = (p.upload(shares)?, p.upload(masks)?, p.upload(unmask0)?, p.upload(unmask1)?, p.upload(unmask2)?);
    let (o0, o1, o2) = (p.alloc(len * SCZ_FR_BYTES)?, p.alloc(len * SCZ_FR_BYTES)?, p.alloc(len * SCZ_FR_BYTES)?);
    crate::check(p, unsafe {
        scz_c_acc_product_and_share_dev(p.ctx(), dpp, d_s.ptr, d_m.ptr, d_u0.ptr, d_u1.ptr, d_u2.ptr, len, o0.ptr, o1.ptr, o2.ptr)
    })?;
    Ok((frs(&p.download::<u64>(&o0, len * 4)?), frs(&p.download::<u64>(&o1, len * 4)?), frs(&p.download::<u64>(&o2, len * 4)?)))
}
