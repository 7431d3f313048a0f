//! The sumcheck family -- dist-primitive/src/dsumcheck.rs, same names, argument order and return shapes:
//!   sumcheck :6-26, sumcheck_product :28-90, c_sumcheck :92-146, c_sumcheck_product :148-285,
//!   d_sumcheck :287-357, d_sumcheck_product :359-512.
//! The local functions (`sumcheck`, `sumcheck_product`) take the net as an extra first argument: it carries the device.
use crate::elements::{pairs_from_limbs, triples_from_limbs, SczFr};
use crate::net::GpuNet;
use mpc_net::{MPCNetError, MultiplexedStreamID};
use scz_sys::*;
use secret_sharing::pss::PackedSharingParams;

fn log2(v: usize) -> usize {
    v.trailing_zeros() as usize
}

pub fn sumcheck<F: SczFr, Net: GpuNet>(net: &Net, evaluation: &Vec<F>, challenge: &Vec<F>) -> Vec<(F, F)> {
    let p = net.gpu();
    let _g = p.lock();
    let cnt = log2(evaluation.len()) + 1; // n round messages + the final (0, f(r)) (:24)
    let (d_f, d_c) = (p.upload(evaluation).expect("upload"), p.upload(challenge).expect("upload"));
    let d_o = p.alloc(cnt * 64).expect("alloc");
    let rc = unsafe { scz_sumcheck_dev(p.ctx(), d_f.ptr, evaluation.len(), d_c.ptr, d_o.ptr) };
    assert_eq!(rc, SCZ_OK, "scz_sumcheck_dev: {}", p.last_error());
    pairs_from_limbs(&p.download::<[u64; 8]>(&d_o, cnt).expect("download"))
}

pub fn sumcheck_product<F: SczFr, Net: GpuNet>(
    net: &Net,
    evaluation_f: &Vec<F>,
    evaluation_g: &Vec<F>,
    challenge: &Vec<F>,
) -> Vec<(F, F, F)> {
    let p = net.gpu();
    let _g = p.lock();
    let cnt = log2(evaluation_f.len()) + 1; // the last triple is (0, f * g, 0) (:87)
    let (d_f, d_g, d_c) = (p.upload(evaluation_f).expect("upload"), p.upload(evaluation_g).expect("upload"),
                           p.upload(challenge).expect("upload"));
    let d_o = p.alloc(cnt * SCZ_TRIPLE_BYTES).expect("alloc");
    let rc = unsafe { scz_sumcheck_product_dev(p.ctx(), d_f.ptr, d_g.ptr, evaluation_f.len(), d_c.ptr, d_o.ptr) };
    assert_eq!(rc, SCZ_OK, "scz_sumcheck_product_dev: {}", p.last_error());
    triples_from_limbs(&p.download::<[u64; 12]>(&d_o, cnt).expect("download"))
}

pub async fn c_sumcheck<F: SczFr, Net: GpuNet>(
    shares: &Vec<F>,
    challenge: &Vec<F>,
    pp: &PackedSharingParams<F>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<Vec<(F, F)>, MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let dpp = p.pp(pp.l)?;
    let cnt = log2(shares.len()) + log2(pp.l) + 1;
    let (d_f, d_c) = (p.upload(shares)?, p.upload(challenge)?);
    let d_o = p.alloc(cnt * 64)?;
    crate::check(p, unsafe { scz_c_sumcheck_dev(p.ctx(), dpp, d_f.ptr, shares.len(), d_c.ptr, d_o.ptr) })?;
    Ok(pairs_from_limbs(&p.download::<[u64; 8]>(&d_o, cnt)?))
}

pub async fn c_sumcheck_product<F: SczFr, Net: GpuNet>(
    shares_f: &Vec<F>,
    shares_g: &Vec<F>,
    challenge: &Vec<F>,
    pp: &PackedSharingParams<F>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<Vec<(F, F, F)>, MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let dpp = p.pp(pp.l)?;
    // n local rounds (:167-219), 2 x pss2ss (:224-225), log2(l) rounds that re-use challenge[0..log2 l) (:230), (0, f g, 0) (:282)
    let cnt = log2(shares_f.len()) + log2(pp.l) + 1;
    let (d_f, d_g, d_c) = (p.upload(shares_f)?, p.upload(shares_g)?, p.upload(challenge)?);
    let d_o = p.alloc(cnt * SCZ_TRIPLE_BYTES)?;
    crate::check(p, unsafe { scz_c_sumcheck_product_dev(p.ctx(), dpp, d_f.ptr, d_g.ptr, shares_f.len(), d_c.ptr, d_o.ptr) })?;
    Ok(triples_from_limbs(&p.download::<[u64; 12]>(&d_o, cnt)?))
}

pub async fn d_sumcheck<F: SczFr, Net: GpuNet>(
    partial_poly: &Vec<F>,
    challenge: &Vec<F>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<Vec<(F, F)>, MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let cap = log2(partial_poly.len()) + log2(net.n_parties());
    let (d_f, d_c) = (p.upload(partial_poly)?, p.upload(challenge)?);
    let d_o = p.alloc(cap.max(1) * 64)?;
    let mut cnt = 0usize;
    crate::check(p, unsafe { scz_d_sumcheck_dev(p.ctx(), d_f.ptr, partial_poly.len(), d_c.ptr, d_o.ptr, &mut cnt) })?;
    Ok(pairs_from_limbs(&p.download::<[u64; 8]>(&d_o, cnt)?)) // non-leaders: empty Vec, like :352-356
}

pub async fn d_sumcheck_product<F: SczFr, Net: GpuNet>(
    partial_f: &Vec<F>,
    partial_g: &Vec<F>,
    challenge: &Vec<F>,
    net: &Net,
    _sid: MultiplexedStreamID,
) -> Result<Vec<(F, F, F)>, MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    // leader: n + log2(N) triples, no trailing final tuple (:452-504); others: empty Vec (:507-509)
    let cap = log2(partial_f.len()) + log2(net.n_parties());
    let (d_f, d_g, d_c) = (p.upload(partial_f)?, p.upload(partial_g)?, p.upload(challenge)?);
    let d_o = p.alloc(cap.max(1) * SCZ_TRIPLE_BYTES)?;
    let mut cnt = 0usize;
    crate::check(p, unsafe { scz_d_sumcheck_product_dev(p.ctx(), d_f.ptr, d_g.ptr, partial_f.len(), d_c.ptr, d_o.ptr, &mut cnt) })?;
    Ok(triples_from_limbs(&p.download::<[u64; 12]>(&d_o, cnt)?))
}
