//! `PackedSharingParams::{pack_from_public, pack_single, unpack, unpack2}` (secret-sharing/src/pss.rs:69-171) on the
//! device, over Fr and over G1 (`G: DomainCoeff<F>`): libscz folds each FFT pair into a small matrix once (csrc/pss.cu).
//! The reference's struct stays the parameter type (`pp.l`, `pp.n`, `pp.t` are read from it); the device twin is
//! cached per party (`GpuParty::pp`).
use crate::elements::{SczFr, SczG1};
use crate::net::GpuNet;
use mpc_net::MPCNetError;
use scz_sys::*;
use secret_sharing::pss::PackedSharingParams;

#[derive(Clone, Copy)]
enum Map {
    PackFromPublic,
    PackSingle,
    Unpack,
    Unpack2,
}

fn apply<N: GpuNet>(net: &N, l: usize, map: Map, kind: i32, input: &[u64], elem_limbs: usize, len_in: usize, len_out: usize)
    -> Result<Vec<u64>, MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let pp = p.pp(l)?;
    assert_eq!(input.len(), len_in * elem_limbs);
    let d_in = p.upload(input)?;
    let d_out = p.alloc(len_out * elem_limbs * 8)?;
    let rc = unsafe {
        match map {
            Map::PackFromPublic => scz_pss_pack_from_public_dev(p.ctx(), pp, kind, d_in.ptr, len_in, 1, d_out.ptr),
            Map::PackSingle => scz_pss_pack_single_dev(p.ctx(), pp, kind, d_in.ptr, 1, d_out.ptr),
            Map::Unpack => scz_pss_unpack_dev(p.ctx(), pp, kind, d_in.ptr, 1, d_out.ptr),
            Map::Unpack2 => scz_pss_unpack2_dev(p.ctx(), pp, kind, d_in.ptr, 1, d_out.ptr),
        }
    };
    crate::check(p, rc)?;
    p.download::<u64>(&d_out, len_out * elem_limbs)
}

fn fr_limbs<F: SczFr>(v: &[F]) -> Vec<u64> {
    let mut out = vec![0u64; v.len() * 4];
    unsafe { core::ptr::copy_nonoverlapping(F::as_raw(v) as *const u64, out.as_mut_ptr(), out.len()) };
    out
}
fn fr_vec<F: SczFr>(l: &[u64]) -> Vec<F> {
    l.chunks_exact(4).map(crate::elements::fr_from::<F>).collect()
}
fn g1_limbs<G: SczG1>(v: &[G]) -> Vec<u64> {
    v.iter().flat_map(|g| g.to_jacobian_limbs()).collect()
}
fn g1_vec<G: SczG1>(l: &[u64]) -> Vec<G> {
    l.chunks_exact(18).map(|c| G::from_jacobian_limbs(c.try_into().unwrap())).collect()
}

/// Device versions of the four maps, with the reference's names.  `secrets.len() <= 2 l` (zero padded, pss.rs:94).
pub trait PackedSharingGpu<F: SczFr> {
    fn pack_from_public_fr<N: GpuNet>(&self, net: &N, secrets: Vec<F>) -> Result<Vec<F>, MPCNetError>;
    fn pack_single_fr<N: GpuNet>(&self, net: &N, secret: F) -> Result<Vec<F>, MPCNetError>;
    fn unpack_fr<N: GpuNet>(&self, net: &N, shares: Vec<F>) -> Result<Vec<F>, MPCNetError>;
    fn unpack2_fr<N: GpuNet>(&self, net: &N, shares: Vec<F>) -> Result<Vec<F>, MPCNetError>;
    fn pack_from_public_g1<G: SczG1, N: GpuNet>(&self, net: &N, secrets: Vec<G>) -> Result<Vec<G>, MPCNetError>;
    fn pack_single_g1<G: SczG1, N: GpuNet>(&self, net: &N, secret: G) -> Result<Vec<G>, MPCNetError>;
    fn unpack_g1<G: SczG1, N: GpuNet>(&self, net: &N, shares: Vec<G>) -> Result<Vec<G>, MPCNetError>;
    fn unpack2_g1<G: SczG1, N: GpuNet>(&self, net: &N, shares: Vec<G>) -> Result<Vec<G>, MPCNetError>;
}
impl<F: SczFr> PackedSharingGpu<F> for PackedSharingParams<F> {
    fn pack_from_public_fr<N: GpuNet>(&self, net: &N, secrets: Vec<F>) -> Result<Vec<F>, MPCNetError> {
        Ok(fr_vec(&apply(net, self.l, Map::PackFromPublic, 0, &fr_limbs(&secrets), 4, secrets.len(), self.n)?))
    }
    fn pack_single_fr<N: GpuNet>(&self, net: &N, secret: F) -> Result<Vec<F>, MPCNetError> {
        Ok(fr_vec(&apply(net, self.l, Map::PackSingle, 0, &fr_limbs(&[secret]), 4, 1, self.n)?))
    }
    fn unpack_fr<N: GpuNet>(&self, net: &N, shares: Vec<F>) -> Result<Vec<F>, MPCNetError> {
        assert_eq!(shares.len(), self.n);
        Ok(fr_vec(&apply(net, self.l, Map::Unpack, 0, &fr_limbs(&shares), 4, self.n, self.l)?))
    }
    fn unpack2_fr<N: GpuNet>(&self, net: &N, shares: Vec<F>) -> Result<Vec<F>, MPCNetError> {
        assert_eq!(shares.len(), self.n);
        Ok(fr_vec(&apply(net, self.l, Map::Unpack2, 0, &fr_limbs(&shares), 4, self.n, self.l)?))
    }
    fn pack_from_public_g1<G: SczG1, N: GpuNet>(&self, net: &N, secrets: Vec<G>) -> Result<Vec<G>, MPCNetError> {
        Ok(g1_vec(&apply(net, self.l, Map::PackFromPublic, 1, &g1_limbs(&secrets), 18, secrets.len(), self.n)?))
    }
    fn pack_single_g1<G: SczG1, N: GpuNet>(&self, net: &N, secret: G) -> Result<Vec<G>, MPCNetError> {
        Ok(g1_vec(&apply(net, self.l, Map::PackSingle, 1, &g1_limbs(&[secret]), 18, 1, self.n)?))
    }
    fn unpack_g1<G: SczG1, N: GpuNet>(&self, net: &N, shares: Vec<G>) -> Result<Vec<G>, MPCNetError> {
        assert_eq!(shares.len(), self.n);
        Ok(g1_vec(&apply(net, self.l, Map::Unpack, 1, &g1_limbs(&shares), 18, self.n, self.l)?))
    }
    fn unpack2_g1<G: SczG1, N: GpuNet>(&self, net: &N, shares: Vec<G>) -> Result<Vec<G>, MPCNetError> {
        assert_eq!(shares.len(), self.n);
        Ok(g1_vec(&apply(net, self.l, Map::Unpack2, 1, &g1_limbs(&shares), 18, self.n, self.l)?))
    }
}
