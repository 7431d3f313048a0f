//! Multilinear KZG -- dist-primitive/src/dpoly_comm.rs:236-464 with the same method names and argument order.
//! `PolynomialCommitment` here is the DEVICE-resident SRS (`scz_srs`): built once from the reference's
//! `powers_of_g` (after `mature()`, :141-150) and reused by every commit / open, optionally with fixed-base tables.
use crate::elements::{fr_from, pack_affine, SczFr, SczG1};
use crate::net::{GpuNet, GpuParty};
use ark_bls12_381::{Bls12_381, Fr, G1Affine, G1Projective};
use core::ffi::c_void;
use mpc_net::{MPCNetError, MultiplexedStreamID};
use scz_sys::*;
use secret_sharing::pss::PackedSharingParams;
use std::ptr;

pub struct PolynomialCommitment<'a> {
    party: &'a GpuParty,
    srs: *mut SczSrs,
    levels: usize,
    _e: core::marker::PhantomData<Bls12_381>,
}
unsafe impl Send for PolynomialCommitment<'_> {}
unsafe impl Sync for PolynomialCommitment<'_> {}
impl Drop for PolynomialCommitment<'_> {
    fn drop(&mut self) {
        unsafe { scz_srs_free(self.srs) };
    }
}

fn g1s(raw: &[[u64; 18]]) -> Vec<G1Projective> {
    raw.iter().map(G1Projective::from_jacobian_limbs).collect()
}

impl<'a> PolynomialCommitment<'a> {
    /// `powers_of_g` as the reference holds it after `mature()` (:141-150): level i = the affine bases of 2^i points
    /// (`new`, :37-67), max(1, 2^i / l) (`new_single`, :197-219) or 2^i (`new_random`, :220-233)
    pub fn from_powers_of_g<Net: GpuNet>(net: &'a Net, powers_of_g: &[Vec<G1Affine>]) -> Result<Self, MPCNetError> {
        let p = net.gpu();
        let _g = p.lock();
        let packed: Vec<_> = powers_of_g.iter().map(|l| pack_affine(l)).collect();
        let ptrs: Vec<*const c_void> = packed.iter().map(|l| l.xy.as_ptr() as *const c_void).collect();
        let lens: Vec<usize> = packed.iter().map(|l| l.xy.len()).collect();
        let mut srs: *mut SczSrs = ptr::null_mut();
        crate::check(p, unsafe { scz_srs_from_host_levels(p.ctx(), ptrs.len(), ptrs.as_ptr(), lens.as_ptr(), &mut srs) })?;
        Ok(Self { party: p, srs, levels: powers_of_g.len(), _e: Default::default() })
    }
    /// `PolynomialCommitmentCub::new(g, g2, s).mature()` (:37-67, :141-150): the real SRS with trapdoor `s`, on the device
    pub fn new<Net: GpuNet>(net: &'a Net, g: G1Projective, s: &[Fr]) -> Result<Self, MPCNetError> {
        let p = net.gpu();
        let _g = p.lock();
        let (d_g, d_s) = (p.upload(&[g.to_jacobian_limbs()])?, p.upload(s)?);
        let mut srs: *mut SczSrs = ptr::null_mut();
        crate::check(p, unsafe { scz_srs_new_dev(p.ctx(), d_g.ptr, d_s.ptr, s.len(), &mut srs) })?;
        Ok(Self { party: p, srs, levels: s.len() + 1, _e: Default::default() })
    }
    /// `to_packed` (:164-194): party `party`'s PSS share of this SRS
    pub fn to_packed(&self, pp: &PackedSharingParams<Fr>, party: u32) -> Result<Self, MPCNetError> {
        let p = self.party;
        let _g = p.lock();
        let mut out: *mut SczSrs = ptr::null_mut();
        crate::check(p, unsafe { scz_srs_to_packed_dev(p.ctx(), self.srs, p.pp(pp.l)?, party, &mut out) })?;
        Ok(Self { party: p, srs: out, levels: self.levels, _e: Default::default() })
    }
    /// window multiples of every level beside the points (csrc/srs.cu): same results, ~20 % fewer bucket additions
    pub fn precompute(self) -> Result<Self, MPCNetError> {
        let _g = self.party.lock();
        crate::check(self.party, unsafe { scz_srs_precompute(self.party.ctx(), self.srs) })?;
        drop(_g);
        Ok(self)
    }
    pub fn raw(&self) -> *const SczSrs {
        self.srs
    }
    pub fn levels(&self) -> usize {
        self.levels
    }

    /// :237-243 (= d_local_commit :269-275).  The asserts of :239-240 stay panics (crate::check).
    pub fn commit(&self, peval: &Vec<Fr>) -> G1Projective {
        let p = self.party;
        let _g = p.lock();
        let d_p = p.upload(peval).expect("upload");
        let d_o = p.alloc(SCZ_G1_JAC_BYTES).expect("alloc");
        crate::check(p, unsafe { scz_commit_dev(p.ctx(), self.srs, d_p.ptr, peval.len(), d_o.ptr) }).expect("commit");
        g1s(&p.download::<[u64; 18]>(&d_o, 1).expect("download"))[0]
    }
    pub fn d_local_commit(&self, peval: &Vec<Fr>) -> G1Projective {
        self.commit(peval)
    }
    /// :244-267: level = log2(len * l); ONE d_msm over the batch
    pub async fn c_commit<Net: GpuNet>(
        &self,
        pevals: &Vec<Vec<Fr>>,
        pp: &PackedSharingParams<Fr>,
        net: &Net,
        _sid: MultiplexedStreamID,
    ) -> Result<Vec<G1Projective>, MPCNetError> {
        let p = net.gpu();
        let _g = p.lock();
        let bufs = pevals.iter().map(|v| p.upload(v)).collect::<Result<Vec<_>, _>>()?;
        let ptrs: Vec<*const c_void> = bufs.iter().map(|b| b.ptr as *const c_void).collect();
        let lens: Vec<usize> = pevals.iter().map(|v| v.len()).collect();
        let d_o = p.alloc(pevals.len().max(1) * SCZ_G1_JAC_BYTES)?;
        crate::check(p, unsafe { scz_c_commit_dev(p.ctx(), self.srs, p.pp(pp.l)?, ptrs.as_ptr(), lens.as_ptr(), pevals.len(), d_o.ptr) })?;
        Ok(g1s(&p.download::<[u64; 18]>(&d_o, pevals.len())?))
    }
    /// :276-297: local commit, the leader sums the N commitments, every party receives the sum
    pub async fn d_commit<Net: GpuNet>(&self, peval: &Vec<Fr>, net: &Net, _sid: MultiplexedStreamID) -> Result<G1Projective, MPCNetError> {
        let p = net.gpu();
        let _g = p.lock();
        let d_p = p.upload(peval)?;
        let d_o = p.alloc(SCZ_G1_JAC_BYTES)?;
        crate::check(p, unsafe { scz_d_commit_dev(p.ctx(), self.srs, d_p.ptr, peval.len(), d_o.ptr) })?;
        Ok(g1s(&p.download::<[u64; 18]>(&d_o, 1)?)[0])
    }
    /// :299-325 (= d_local_open :327-353): value + n proofs
    pub fn open(&self, peval: &Vec<Fr>, point: &[Fr]) -> (Fr, Vec<G1Projective>) {
        let p = self.party;
        let _g = p.lock();
        let n = peval.len().trailing_zeros() as usize;
        assert_eq!(peval.len(), 1usize << n); // :306
        let (d_p, d_u) = (p.upload(peval).expect("upload"), p.upload(point).expect("upload"));
        let (d_v, d_o) = (p.alloc(SCZ_FR_BYTES).expect("alloc"), p.alloc(n.max(1) * SCZ_G1_JAC_BYTES).expect("alloc"));
        crate::check(p, unsafe { scz_open_dev(p.ctx(), self.srs, d_p.ptr, peval.len(), d_u.ptr, d_v.ptr, d_o.ptr) }).expect("open");
        (fr_from::<Fr>(&p.download::<u64>(&d_v, 4).expect("download")), g1s(&p.download::<[u64; 18]>(&d_o, n).expect("download")))
    }
    pub fn d_local_open(&self, peval: &Vec<Fr>, point: &[Fr]) -> (Fr, Vec<G1Projective>) {
        self.open(peval, point)
    }
    /// :355-398: leader gets value + log2(N) root proofs ++ n summed proofs, workers (0, [])
    pub async fn d_open<Net: GpuNet>(
        &self,
        peval: &Vec<Fr>,
        point: &Vec<Fr>,
        net: &Net,
        _sid: MultiplexedStreamID,
    ) -> Result<(Fr, Vec<G1Projective>), MPCNetError> {
        let p = net.gpu();
        let _g = p.lock();
        let cap = peval.len().trailing_zeros() as usize + net.n_parties().trailing_zeros() as usize;
        let (d_p, d_u) = (p.upload(peval)?, p.upload(point)?);
        let (d_v, d_o) = (p.alloc(SCZ_FR_BYTES)?, p.alloc(cap.max(1) * SCZ_G1_JAC_BYTES)?);
        let mut cnt = 0usize;
        crate::check(p, unsafe {
            scz_d_open_dev(p.ctx(), self.srs, d_p.ptr, peval.len(), d_u.ptr, point.len(), d_v.ptr, d_o.ptr, &mut cnt)
        })?;
        Ok((fr_from::<Fr>(&p.download::<u64>(&d_v, 4)?), g1s(&p.download::<[u64; 18]>(&d_o, cnt)?)))
    }
    /// :401-464: folds first, ONE batched c_commit of all quotients (:436), pss2ss(r) (:439), log2(l) tail rounds that
    /// index point[i] from 0 (:452)
    pub async fn c_open<Net: GpuNet>(
        &self,
        peval: &Vec<Fr>,
        point: &Vec<Fr>,
        pp: &PackedSharingParams<Fr>,
        net: &Net,
        _sid: MultiplexedStreamID,
    ) -> Result<(Fr, Vec<G1Projective>), MPCNetError> {
        let p = net.gpu();
        let _g = p.lock();
        let cnt = peval.len().trailing_zeros() as usize + pp.l.trailing_zeros() as usize;
        let (d_p, d_u) = (p.upload(peval)?, p.upload(point)?);
        let (d_v, d_o) = (p.alloc(SCZ_FR_BYTES)?, p.alloc(cnt.max(1) * SCZ_G1_JAC_BYTES)?);
        crate::check(p, unsafe { scz_c_open_dev(p.ctx(), self.srs, p.pp(pp.l)?, d_p.ptr, peval.len(), d_u.ptr, d_v.ptr, d_o.ptr) })?;
        Ok((fr_from::<Fr>(&p.download::<u64>(&d_v, 4)?), g1s(&p.download::<[u64; 18]>(&d_o, cnt)?)))
    }
}
#[allow(dead_code)]
fn _bounds<F: SczFr, G: SczG1>() {}
