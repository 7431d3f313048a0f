//! Raw binding of `include/scz.h`.  The `extern "C"` block (`ffi.rs`) is GENERATED from the header by
//! `tools/gen_scz_sys.py` (one declaration per prototype; `tests/test_rust_shim.py` keeps the two in step); the opaque
//! handles, the plain-data structs and the callback table below are written by hand and mirror the header's typedefs
//! field by field.  No Rust toolchain exists in the authoring image: this crate has not been compiled there.
#![allow(non_camel_case_types)]
use core::ffi::c_void;

mod ffi;
pub use ffi::*;

pub const SCZ_OK: i32 = 0;
pub const SCZ_ERR_BAD_ARG: i32 = -1;
pub const SCZ_ERR_LEN_MISMATCH: i32 = -2; // ark-ec msm Err(min len) -> the reference unwrap()s: dmsm.rs:23
pub const SCZ_ERR_NOT_POW2: i32 = -3; // dpoly_comm.rs:240,255 asserts
pub const SCZ_ERR_LEVEL_OOB: i32 = -4; // dpoly_comm.rs:239,254 asserts
pub const SCZ_ERR_CUDA: i32 = -5;
pub const SCZ_ERR_NET: i32 = -6; // MPCNetError (mpc-net/src/lib.rs:14-26)
pub const SCZ_ERR_NOMEM: i32 = -7;

pub const SCZ_FR_BYTES: usize = 32;
pub const SCZ_G1_AFFINE_BYTES: usize = 96;
pub const SCZ_G1_JAC_BYTES: usize = 144;
pub const SCZ_TRIPLE_BYTES: usize = 96;
pub const SCZ_G2_AFFINE_BYTES: usize = 192; // x | y, each Fq2 = c0 | c1
pub const SCZ_G2_JAC_BYTES: usize = 288; // X | Y | Z = Projective<g2::Config>
pub const SCZ_NCCL_UID_BYTES: usize = 128;
pub const SCZ_STATUS_DIV_BY_ZERO: u32 = 1;

pub const SCZ_HP_GATE_PROOF: u32 = 0;
pub const SCZ_HP_GATE_COMMIT: u32 = 1;
pub const SCZ_HP_WIRING_PROOF: u32 = 2;
pub const SCZ_HP_WIRING_COMMIT: u32 = 3;
pub const SCZ_HP_WIRING_OPEN: u32 = 4;

macro_rules! opaque {
    ($($name:ident),*) => { $( #[repr(C)] pub struct $name { _private: [u8; 0] } )* };
}
opaque!(SczCtx, SczPp, SczSrs, SczNcclHub);

/// `scz_net_vtable`: the MPCSerializeNet seam for hosts that keep their own transport
/// (dist-primitive/src/utils/serializing_net.rs:8-142).  Payloads are DEVICE buffers.
pub type SczColl = unsafe extern "C" fn(
    user: *mut c_void, d_send: *const c_void, d_recv: *mut c_void, bytes: usize, wire_bytes: usize, stream: *mut c_void,
) -> i32;
pub type SczRooted = unsafe extern "C" fn(
    user: *mut c_void, root: u32, d_send: *const c_void, d_recv: *mut c_void, bytes: usize, wire_bytes: usize,
    stream: *mut c_void,
) -> i32;
pub type SczSync = unsafe extern "C" fn(user: *mut c_void, stream: *mut c_void) -> i32;
#[repr(C)]
pub struct SczNetVtable {
    pub user: *mut c_void,
    pub gather: SczColl,     // worker_send_or_leader_receive_element (:11-39)
    pub scatter: SczColl,    // worker_receive_or_leader_send_element (:76-96)
    pub all_gather: SczColl, // the N hub rounds of hyperplonk/src/dhyperplonk.rs:271-294
    pub sync: Option<SczSync>, // MPCNet::sync (mpc-net/src/lib.rs:275-286)
    pub gather_to: Option<SczRooted>, // dynamic_ variants (:41-74)
    pub scatter_from: Option<SczRooted>, // (:98-126)
}

/// `scz_hp_pk`: the fields of PackedProvingParameters (hyperplonk/src/dhyperplonk.rs:22-62) `dhyperplonk` reads, as
/// device pointers, in the header's order; plus the three vectors the reference draws from entropy (:188-190).
#[repr(C)]
pub struct SczHpPk {
    pub v: *const c_void,
    pub a_evals: *const c_void,
    pub b_evals: *const c_void,
    pub c_evals: *const c_void,
    pub i: *const c_void,
    pub s1: *const c_void,
    pub s2: *const c_void,
    pub i_p: *const c_void,
    pub s1_p: *const c_void,
    pub s2_p: *const c_void,
    pub ssigma_p: *const c_void,
    pub sid_p: *const c_void,
    pub eq: *const c_void,
    pub eq_r1_p: *const c_void,
    pub eq_r2_p: *const c_void,
    pub challenge: *const c_void,
    pub challenge_r1: *const c_void,
    pub challenge_r2: *const c_void,
    pub alpha_beta: *const c_void,
    pub c_commitment: *const SczSrs,
    pub d_commitment: *const SczSrs,
    pub local_s_p: *const c_void,
    pub local_s: *const c_void,
    pub eq_leader: *const c_void,
}

/// `scz_hp_item`: one entry of the prover's return value in the reference's push order (dhyperplonk.rs:567-570)
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct SczHpItem {
    pub kind: u32,
    pub triples_off: u32,
    pub triples_cnt: u32,
    pub points_off: u32,
    pub points_cnt: u32,
    pub value_off: u32,
    pub value_cnt: u32,
}

/// `scz_local_pk`: local_hyperplonk's tables (hyperplonk/src/hyperplonk.rs:15-160)
#[repr(C)]
pub struct SczLocalPk {
    pub m: *const c_void,
    pub a_evals: *const c_void,
    pub b_evals: *const c_void,
    pub c_evals: *const c_void,
    pub input: *const c_void,
    pub q1: *const c_void,
    pub q2: *const c_void,
    pub ssigma: *const c_void,
    pub sid: *const c_void,
    pub eq: *const c_void,
    pub eq_p2: *const c_void,
    pub challenge: *const c_void,
    pub challengep2: *const c_void,
    pub alpha_beta: *const c_void,
    pub commitment: *const SczSrs,
}

/// `scz_cperm_pk`: cpermcheck's tables (hyperplonk/src/dhyperplonk.rs:1249-1385)
#[repr(C)]
pub struct SczCpermPk {
    pub v: *const c_void,
    pub sid: *const c_void,
    pub ssigma: *const c_void,
    pub eq_r1: *const c_void,
    pub mask: *const c_void,
    pub unmask0: *const c_void,
    pub unmask1: *const c_void,
    pub unmask2: *const c_void,
    pub challenge_r1: *const c_void,
    pub alpha_beta: *const c_void,
    pub c_commitment: *const SczSrs,
}
