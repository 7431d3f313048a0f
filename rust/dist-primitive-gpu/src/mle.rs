//! `fix_variable` -- dist-primitive/src/mle.rs:88-104: folds the top min(points.len(), log2 len) variables.
use crate::elements::{fr_from, SczFr};
use crate::net::GpuNet;
use scz_sys::*;

pub fn fix_variable<F: SczFr, Net: GpuNet>(net: &Net, evaluations: &Vec<F>, points: &Vec<F>) -> Vec<F> {
    let p = net.gpu();
    let _g = p.lock();
    let n = evaluations.len().trailing_zeros() as usize;
    let folded = points.len().min(n);
    let d_e = p.upload(evaluations).expect("upload");
    let d_p = p.upload(points).expect("upload");
    let out_len = evaluations.len() >> folded;
    let d_o = p.alloc(out_len * SCZ_FR_BYTES).expect("alloc");
    let rc = unsafe { scz_fix_variable_dev(p.ctx(), d_e.ptr, evaluations.len(), d_p.ptr, points.len(), d_o.ptr) };
    assert_eq!(rc, SCZ_OK, "scz_fix_variable_dev: {}", p.last_error());
    p.download::<u64>(&d_o, out_len * 4).expect("download").chunks_exact(4).map(fr_from::<F>).collect()
}
