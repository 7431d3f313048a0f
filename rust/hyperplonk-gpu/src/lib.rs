//! The prover entries of `hyperplonk/src/dhyperplonk.rs` -- `dhyperplonk` (:159-571), `dhyperplonk_data_parallel`
//! (:573-960), `dpermcheck` (:962-1247), `cpermcheck` (:1249-1385) -- and the monolithic `local_hyperplonk`
//! (hyperplonk/src/hyperplonk.rs:15-160), same names, same nested return tuples.
//!
//! Two differences, both forced by the device: (1) the proving parameters are uploaded ONCE
//! (`GpuProvingParameters::upload`, like `PackedProvingParameters::new` runs once before the timed call in
//! bench_hyperplonk.rs:44-46) and stay resident; (2) the three vectors the reference draws from OS entropy INSIDE
//! `dhyperplonk` (:188-190) are explicit inputs (`InjectedEntropy`), otherwise no two runs could be compared.
use ark_bls12_381::{Bls12_381, Fr, G1Projective};
use core::ffi::c_void;
use dist_primitive_gpu::dpoly_comm::PolynomialCommitment;
use dist_primitive_gpu::elements::{fr_from, triples_from_limbs, SczG1};
use dist_primitive_gpu::net::{DevBuf, GpuNet, GpuParty};
use hyperplonk::dhyperplonk::PackedProvingParameters;
use mpc_net::{MPCNetError, MultiplexedStreamID};
use scz_sys::*;
use secret_sharing::pss::PackedSharingParams;

type Triples = Vec<Vec<(Fr, Fr, Fr)>>;
/// the reference's return type (:165-186)
pub type Proof = ((Triples, Vec<(G1Projective, (Fr, Vec<G1Projective>))>), (Triples, Vec<G1Projective>, Vec<(Fr, Vec<G1Projective>)>));

/// `local_s_p`, `local_s`, `eq` of dhyperplonk.rs:188-190
pub struct InjectedEntropy {
    pub local_s_p: Vec<Fr>,
    pub local_s: Vec<Fr>,
    pub eq_leader: Vec<Fr>,
}

/// The tables of `PackedProvingParameters` that the provers read, resident in HBM, plus the two device SRSs
pub struct GpuProvingParameters<'a> {
    bufs: Vec<DevBuf<'a>>, // in scz_hp_pk order: V .. alpha_beta, then local_s_p, local_s, eq_leader
    c_commitment: PolynomialCommitment<'a>,
    d_commitment: PolynomialCommitment<'a>,
}
impl<'a> GpuProvingParameters<'a> {
    pub fn upload<Net: GpuNet>(
        net: &'a Net,
        pk: &PackedProvingParameters<Bls12_381>,
        c_powers_of_g: &[Vec<ark_bls12_381::G1Affine>],
        d_powers_of_g: &[Vec<ark_bls12_381::G1Affine>],
        entropy: &InjectedEntropy,
        fixed_base_tables: bool,
    ) -> Result<Self, MPCNetError> {
        let p: &GpuParty = net.gpu();
        let alpha_beta = vec![pk.alpha, pk.beta];
        let tables: [&Vec<Fr>; 22] = [
            &pk.V, &pk.a_evals, &pk.b_evals, &pk.c_evals, &pk.I, &pk.S1, &pk.S2, &pk.I_p, &pk.S1_p, &pk.S2_p, &pk.ssigma_p,
            &pk.sid_p, &pk.eq, &pk.eq_r1_p, &pk.eq_r2_p, &pk.challenge, &pk.challenge_r1, &pk.challenge_r2, &alpha_beta,
            &entropy.local_s_p, &entropy.local_s, &entropy.eq_leader,
        ];
        let bufs = {
            let _g = p.lock();
            tables.iter().map(|t| p.upload(t)).collect::<Result<Vec<_>, _>>()?
        };
        let (mut c, mut d) = (PolynomialCommitment::from_powers_of_g(net, c_powers_of_g)?, PolynomialCommitment::from_powers_of_g(net, d_powers_of_g)?);
        if fixed_base_tables {
            c = c.precompute()?;
            d = d.precompute()?;
        }
        Ok(Self { bufs, c_commitment: c, d_commitment: d })
    }
    fn raw(&self) -> SczHpPk {
        let b = |i: usize| self.bufs[i].ptr as *const c_void;
        SczHpPk {
            v: b(0), a_evals: b(1), b_evals: b(2), c_evals: b(3), i: b(4), s1: b(5), s2: b(6), i_p: b(7), s1_p: b(8), s2_p: b(9),
            ssigma_p: b(10), sid_p: b(11), eq: b(12), eq_r1_p: b(13), eq_r2_p: b(14), challenge: b(15), challenge_r1: b(16),
            challenge_r2: b(17), alpha_beta: b(18), c_commitment: self.c_commitment.raw(), d_commitment: self.d_commitment.raw(),
            local_s_p: b(19), local_s: b(20), eq_leader: b(21),
        }
    }
}

type ProverFn = unsafe extern "C" fn(*mut SczCtx, usize, *const SczHpPk, *const SczPp, *mut c_void, usize, *mut c_void, usize,
                                     *mut c_void, usize, *mut SczHpItem, usize, *mut usize) -> i32;

async fn run<Net: GpuNet>(f: ProverFn, n: usize, pk: &GpuProvingParameters<'_>, pp: &PackedSharingParams<Fr>, net: &Net) -> Result<Proof, MPCNetError> {
    net.sync(MultiplexedStreamID::Zero).await?; // dhyperplonk.rs:193
    let p = net.gpu();
    let _g = p.lock();
    let (mut nt, mut np, mut nv, mut ni) = (0usize, 0usize, 0usize, 0usize);
    unsafe { scz_dhyperplonk_sizes(n, pp.l, net.n_parties(), &mut nt, &mut np, &mut nv, &mut ni) };
    let (tri, pts, val) = (p.alloc(nt * SCZ_TRIPLE_BYTES)?, p.alloc(np * SCZ_G1_JAC_BYTES)?, p.alloc(nv * SCZ_FR_BYTES)?);
    let mut items = vec![SczHpItem::default(); ni];
    let mut cnt = 0usize;
    let raw = pk.raw();
    let rc = unsafe { f(p.ctx(), n, &raw, p.pp(pp.l)?, tri.ptr, nt, pts.ptr, np, val.ptr, nv, items.as_mut_ptr(), ni, &mut cnt) };
    if rc != SCZ_OK {
        return Err(MPCNetError::Generic(p.last_error()));
    }
    let tri = triples_from_limbs::<Fr>(&p.download::<[u64; 12]>(&tri, nt)?);
    let pts: Vec<G1Projective> = p.download::<[u64; 18]>(&pts, np)?.iter().map(G1Projective::from_jacobian_limbs).collect();
    let val: Vec<Fr> = p.download::<u64>(&val, nv * 4)?.chunks_exact(4).map(fr_from::<Fr>).collect();
    p.panic_on_status(); // h = num / den (:338-339): arkworks panics on den = 0
    if net.is_leader() {
        println!("Comm: {:?}", net.get_comm()); // :563-565
    }
    Ok(rebuild_nested(&items[..cnt], &tri, &pts, &val))
}

/// the item table names every slice of the three arenas in the reference's push order (include/scz.h, scz_hp_item)
fn rebuild_nested(items: &[SczHpItem], tri: &[(Fr, Fr, Fr)], pts: &[G1Projective], val: &[Fr]) -> Proof {
    let (mut gp, mut gc, mut wp, mut wc, mut wo) = (vec![], vec![], vec![], vec![], vec![]);
    for it in items {
        let t = tri[it.triples_off as usize..(it.triples_off + it.triples_cnt) as usize].to_vec();
        let p = &pts[it.points_off as usize..(it.points_off + it.points_cnt) as usize];
        let v = if it.value_cnt > 0 { val[it.value_off as usize] } else { Fr::from(0u64) };
        match it.kind {
            SCZ_HP_GATE_PROOF => gp.push(t),
            SCZ_HP_GATE_COMMIT => gc.push((p[0], (v, p[1..].to_vec()))),
            SCZ_HP_WIRING_PROOF => wp.push(t),
            SCZ_HP_WIRING_COMMIT => wc.push(p[0]),
            _ => wo.push((v, p.to_vec())),
        }
    }
    ((gp, gc), (wp, wc, wo))
}

/// dhyperplonk.rs:159-571
pub async fn dhyperplonk<Net: GpuNet>(n: usize, pk: &GpuProvingParameters<'_>, pp: &PackedSharingParams<Fr>, net: &Net,
                                      _sid: MultiplexedStreamID) -> Result<Proof, MPCNetError> {
    run(scz_dhyperplonk_dev, n, pk, pp, net).await
}
/// dhyperplonk.rs:573-960: `entropy.local_s` holds the whole s (4 * 2^n / l entries, :603), no exchange
pub async fn dhyperplonk_data_parallel<Net: GpuNet>(n: usize, pk: &GpuProvingParameters<'_>, pp: &PackedSharingParams<Fr>, net: &Net,
                                                    _sid: MultiplexedStreamID) -> Result<Proof, MPCNetError> {
    run(scz_dhyperplonk_data_parallel_dev, n, pk, pp, net).await
}
/// dhyperplonk.rs:962-1247: the wiring identity alone (the gate half of the tuple comes back empty)
pub async fn dpermcheck<Net: GpuNet>(n: usize, pk: &GpuProvingParameters<'_>, pp: &PackedSharingParams<Fr>, net: &Net,
                                     _sid: MultiplexedStreamID) -> Result<Proof, MPCNetError> {
    run(scz_dpermcheck_dev, n, pk, pp, net).await
}

/// cpermcheck (dhyperplonk.rs:1249-1385): the masked PSS prodcheck, the paper's baseline
pub async fn cpermcheck<Net: GpuNet>(n: usize, pk: &PackedProvingParameters<Bls12_381>, c_commitment: &PolynomialCommitment<'_>,
                                     pp: &PackedSharingParams<Fr>, net: &Net, _sid: MultiplexedStreamID) -> Result<Proof, MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let alpha_beta = vec![pk.alpha, pk.beta];
    let tabs: [&Vec<Fr>; 10] = [&pk.V, &pk.sid, &pk.ssigma, &pk.eq_r1, &pk.mask, &pk.unmask0, &pk.unmask1, &pk.unmask2, &pk.challenge_r1, &alpha_beta];
    let bufs = tabs.iter().map(|t| p.upload(t)).collect::<Result<Vec<_>, _>>()?;
    let b = |i: usize| bufs[i].ptr as *const c_void;
    let raw = SczCpermPk { v: b(0), sid: b(1), ssigma: b(2), eq_r1: b(3), mask: b(4), unmask0: b(5), unmask1: b(6), unmask2: b(7),
                           challenge_r1: b(8), alpha_beta: b(9), c_commitment: c_commitment.raw() };
    let (mut nt, mut np, mut nv, mut ni) = (0usize, 0usize, 0usize, 0usize);
    unsafe { scz_dhyperplonk_sizes(n, pp.l, net.n_parties(), &mut nt, &mut np, &mut nv, &mut ni) };
    let (tri, pts, val) = (p.alloc(nt * SCZ_TRIPLE_BYTES)?, p.alloc(np * SCZ_G1_JAC_BYTES)?, p.alloc(nv * SCZ_FR_BYTES)?);
    let mut items = vec![SczHpItem::default(); ni];
    let mut cnt = 0usize;
    let rc = unsafe { scz_cpermcheck_dev(p.ctx(), n, &raw, p.pp(pp.l)?, tri.ptr, nt, pts.ptr, np, val.ptr, nv, items.as_mut_ptr(), ni, &mut cnt) };
    if rc != SCZ_OK {
        return Err(MPCNetError::Generic(p.last_error()));
    }
    let tri = triples_from_limbs::<Fr>(&p.download::<[u64; 12]>(&tri, nt)?);
    let pts: Vec<G1Projective> = p.download::<[u64; 18]>(&pts, np)?.iter().map(G1Projective::from_jacobian_limbs).collect();
    let val: Vec<Fr> = p.download::<u64>(&val, nv * 4)?.chunks_exact(4).map(fr_from::<Fr>).collect();
    p.panic_on_status();
    Ok(rebuild_nested(&items[..cnt], &tri, &pts, &val))
}

/// local_hyperplonk (hyperplonk/src/hyperplonk.rs:15-160): plain tables, one prover, the "Local HyperPlonk" baseline
pub struct LocalTables<'t> {
    pub m: &'t Vec<Fr>, pub a_evals: &'t Vec<Fr>, pub b_evals: &'t Vec<Fr>, pub c_evals: &'t Vec<Fr>, pub input: &'t Vec<Fr>,
    pub q1: &'t Vec<Fr>, pub q2: &'t Vec<Fr>, pub ssigma: &'t Vec<Fr>, pub sid: &'t Vec<Fr>, pub eq: &'t Vec<Fr>, pub eq_p2: &'t Vec<Fr>,
    pub challenge: &'t Vec<Fr>, pub challengep2: &'t Vec<Fr>, pub alpha: Fr, pub beta: Fr,
}
pub fn local_hyperplonk<Net: GpuNet>(net: &Net, n: usize, t: &LocalTables<'_>, commitment: &PolynomialCommitment<'_>) -> Result<Proof, MPCNetError> {
    let p = net.gpu();
    let _g = p.lock();
    let alpha_beta = vec![t.alpha, t.beta];
    let tabs: [&Vec<Fr>; 14] = [t.m, t.a_evals, t.b_evals, t.c_evals, t.input, t.q1, t.q2, t.ssigma, t.sid, t.eq, t.eq_p2, t.challenge,
                                t.challengep2, &alpha_beta];
    let bufs = tabs.iter().map(|v| p.upload(v)).collect::<Result<Vec<_>, _>>()?;
    let b = |i: usize| bufs[i].ptr as *const c_void;
    let raw = SczLocalPk { m: b(0), a_evals: b(1), b_evals: b(2), c_evals: b(3), input: b(4), q1: b(5), q2: b(6), ssigma: b(7), sid: b(8),
                           eq: b(9), eq_p2: b(10), challenge: b(11), challengep2: b(12), alpha_beta: b(13), commitment: commitment.raw() };
    let (mut nt, mut np, mut nv, mut ni) = (0usize, 0usize, 0usize, 0usize);
    unsafe { scz_dhyperplonk_sizes(n, 1, 8, &mut nt, &mut np, &mut nv, &mut ni) };
    let (tri, pts, val) = (p.alloc(nt * SCZ_TRIPLE_BYTES)?, p.alloc(np * SCZ_G1_JAC_BYTES)?, p.alloc(nv * SCZ_FR_BYTES)?);
    let mut items = vec![SczHpItem::default(); ni];
    let mut cnt = 0usize;
    let rc = unsafe { scz_local_hyperplonk_dev(p.ctx(), n, &raw, tri.ptr, nt, pts.ptr, np, val.ptr, nv, items.as_mut_ptr(), ni, &mut cnt) };
    if rc != SCZ_OK {
        return Err(MPCNetError::Generic(p.last_error()));
    }
    let tri = triples_from_limbs::<Fr>(&p.download::<[u64; 12]>(&tri, nt)?);
    let pts: Vec<G1Projective> = p.download::<[u64; 18]>(&pts, np)?.iter().map(G1Projective::from_jacobian_limbs).collect();
    let val: Vec<Fr> = p.download::<u64>(&val, nv * 4)?.chunks_exact(4).map(fr_from::<Fr>).collect();
    p.panic_on_status();
    Ok(rebuild_nested(&items[..cnt], &tri, &pts, &val))
}
