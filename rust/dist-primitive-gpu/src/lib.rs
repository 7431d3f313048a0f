//! `dist-primitive` on B200: the same public names and signatures as the reference crate
//! (`dist-primitive/src/{dmsm,dsumcheck,dpoly_comm,dacc_product,unpack,degree_reduce,mle}.rs`), every body a thin
//! call into `libscz.so` (`include/scz.h`).  The only type that changes meaning is the net: the reference's functions
//! are generic over `Net: MPCSerializeNet`; here the bound is `Net: GpuNet`, a net that owns this party's `scz_ctx`
//! (leader simulator, or NCCL over NVLink in place of TCP).  Both nets of `net.rs` also implement the reference's
//! `MPCNet`, so code that still moves bytes itself keeps working.
//!
//! libscz implements BLS12-381 only (the curve of the prover path, hyperplonk/examples/hyperplonk.rs:38,48): the
//! generic parameters are kept so that call sites compile unchanged, and sealed to `ark_bls12_381` through
//! `SczFr` / `SczG1`.  Source only: not compiled in the authoring image (no Rust toolchain), see INTEGRATION.md.
pub mod dacc_product;
pub mod degree_reduce;
pub mod dmsm;
pub mod dpoly_comm;
pub mod dsumcheck;
pub mod elements;
pub mod mle;
pub mod net;
pub mod pss;
pub mod unpack;

pub use elements::{SczFr, SczG1};
pub use net::{GpuNet, GpuParty, LeaderGpuNet, NcclGpuNet};

use mpc_net::MPCNetError;

/// Maps a libscz status to the reference's error type (mpc-net/src/lib.rs:14-26).  A base / scalar length mismatch is
/// a PANIC in the reference (`G::msm(..).unwrap()`, dmsm.rs:23) and stays one; so do the power-of-two / level asserts
/// of dpoly_comm.rs:239-240, 254-255.
pub(crate) fn check(party: &GpuParty, rc: i32) -> Result<(), MPCNetError> {
    match rc {
        scz_sys::SCZ_OK => Ok(()),
        scz_sys::SCZ_ERR_LEN_MISMATCH => panic!("called `Result::unwrap()` on an `Err` value: {}", party.last_error()),
        scz_sys::SCZ_ERR_NOT_POW2 | scz_sys::SCZ_ERR_LEVEL_OOB => panic!("assertion failed: {}", party.last_error()),
        scz_sys::SCZ_ERR_NET => Err(MPCNetError::Protocol { err: party.last_error(), party: party.party_id() }),
        scz_sys::SCZ_ERR_BAD_ARG => Err(MPCNetError::BadInput { err: "libscz: bad argument" }),
        _ => Err(MPCNetError::Generic(party.last_error())),
    }
}
