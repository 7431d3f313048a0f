//! The net a GPU party works on.  The reference's only polymorphic seam is `Net: MPCSerializeNet`
//! (dist-primitive/src/utils/serializing_net.rs:266, blanket over `MPCNet`, mpc-net/src/lib.rs:35-61); here a net also
//! OWNS the party's device context, because the collectives of the hot path run on device buffers inside libscz:
//!   * `LeaderGpuNet`  = the reference's build without feature `comm` (serializing_net.rs:144-264): one party, the
//!                       leader sees N clones of its own message; nothing moves
//!   * `NcclGpuNet`    = feature `comm` on one multi-GPU box: one process per GPU, star rounds as grouped
//!                       ncclSend / ncclRecv over NVLink issued by libscz on the ctx stream (csrc/nccl_net.cu)
//! Both implement `MPCNet` (byte-level `send_to` / `recv_from`, counters), so the reference's own helpers
//! (`worker_send_or_leader_receive`, `leader_compute`, `sync`, ...) keep working on top of them.
use async_trait::async_trait;
use bytes::Bytes;
use core::ffi::c_void;
use mpc_net::{MPCNet, MPCNetError, MultiplexedStreamID};
use scz_sys::*;
use std::ffi::CStr;
use std::ptr;
use std::sync::Mutex;

/// One party's `scz_ctx` (+ the `scz_pp` of the packing factor in use).  Calls on a ctx are serialised by the mutex:
/// the reference's party tasks may hop OS threads (mpc-net/src/multi.rs:345-348), libscz sets the device itself.
pub struct GpuParty {
    ctx: *mut SczCtx,
    pp: Mutex<Option<(usize, *mut SczPp)>>,
    party_id: u32,
    n_parties: usize,
    lock: Mutex<()>,
}
unsafe impl Send for GpuParty {}
unsafe impl Sync for GpuParty {}

impl GpuParty {
    pub fn ctx(&self) -> *mut SczCtx {
        self.ctx
    }
    pub fn party_id(&self) -> u32 {
        self.party_id
    }
    pub fn n_parties(&self) -> usize {
        self.n_parties
    }
    pub fn lock(&self) -> std::sync::MutexGuard<'_, ()> {
        self.lock.lock().unwrap()
    }
    pub fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(scz_last_error(self.ctx)).to_string_lossy().into_owned() }
    }
    /// `scz_pp` for packing factor `l` (PackedSharingParams::new, secret-sharing/src/pss.rs:38-65), built once
    pub fn pp(&self, l: usize) -> Result<*const SczPp, MPCNetError> {
        let mut g = self.pp.lock().unwrap();
        if let Some((have, p)) = *g {
            if have == l {
                return Ok(p);
            }
            unsafe { scz_pp_free(p) };
        }
        let mut p: *mut SczPp = ptr::null_mut();
        crate::check(self, unsafe { scz_pp_new(self.ctx, l, &mut p) })?;
        *g = Some((l, p));
        Ok(p)
    }
    // ---- device buffers
    pub fn upload<T>(&self, host: &[T]) -> Result<DevBuf<'_>, MPCNetError> {
        let bytes = std::mem::size_of_val(host);
        let b = self.alloc(bytes)?;
        if bytes > 0 {
            crate::check(self, unsafe { scz_h2d(self.ctx, b.ptr, host.as_ptr() as *const c_void, bytes) })?;
        }
        Ok(b)
    }
    pub fn alloc(&self, bytes: usize) -> Result<DevBuf<'_>, MPCNetError> {
        let mut p: *mut c_void = ptr::null_mut();
        crate::check(self, unsafe { scz_dev_alloc(self.ctx, bytes.max(16), &mut p) })?;
        Ok(DevBuf { party: self, ptr: p, bytes })
    }
    pub fn download<T: Copy + Default>(&self, buf: &DevBuf<'_>, count: usize) -> Result<Vec<T>, MPCNetError> {
        let mut out = vec![T::default(); count];
        let bytes = count * std::mem::size_of::<T>();
        assert!(bytes <= buf.bytes);
        if bytes > 0 {
            crate::check(self, unsafe { scz_d2h(self.ctx, out.as_mut_ptr() as *mut c_void, buf.ptr, bytes) })?;
        }
        Ok(out)
    }
    /// arkworks panics on a division by zero (hyperplonk/src/dhyperplonk.rs:338-339); libscz records it
    pub fn panic_on_status(&self) {
        let mut bits = 0u32;
        unsafe { scz_ctx_take_status(self.ctx, &mut bits) };
        if bits & SCZ_STATUS_DIV_BY_ZERO != 0 {
            panic!("attempt to divide by zero (Field::div)");
        }
    }
    pub fn get_comm(&self) -> (usize, usize) {
        let (mut up, mut down) = (0u64, 0u64);
        unsafe { scz_ctx_get_comm(self.ctx, &mut up, &mut down) };
        (up as usize, down as usize)
    }
}
impl Drop for GpuParty {
    fn drop(&mut self) {
        if let Some((_, p)) = self.pp.lock().unwrap().take() {
            unsafe { scz_pp_free(p) };
        }
        unsafe { scz_ctx_destroy(self.ctx) };
    }
}

pub struct DevBuf<'a> {
    party: &'a GpuParty,
    pub ptr: *mut c_void,
    pub bytes: usize,
}
impl Drop for DevBuf<'_> {
    fn drop(&mut self) {
        unsafe { scz_dev_free(self.party.ctx, self.ptr) };
    }
}

/// A net that owns a GPU party.  The hot-path functions of this crate are generic over it where the reference's are
/// generic over `MPCSerializeNet`.
pub trait GpuNet: MPCNet {
    fn gpu(&self) -> &GpuParty;
}

// ------------------------------------------------------------------------------------------ leader simulator
pub struct LeaderGpuNet {
    party: GpuParty,
}
impl LeaderGpuNet {
    /// `n_parties` = 8 l (pss.rs:39); the single simulated party is the leader
    pub fn new(device: i32, n_parties: usize) -> Result<Self, MPCNetError> {
        let mut ctx: *mut SczCtx = ptr::null_mut();
        let rc = unsafe { scz_ctx_create(device, 0, n_parties as u32, ptr::null(), &mut ctx) };
        if rc != SCZ_OK {
            return Err(MPCNetError::Generic(format!("scz_ctx_create failed ({rc}): no CUDA device? libscz has no CPU path")));
        }
        Ok(Self { party: GpuParty { ctx, pp: Mutex::new(None), party_id: 0, n_parties, lock: Mutex::new(()) } })
    }
}
impl GpuNet for LeaderGpuNet {
    fn gpu(&self) -> &GpuParty {
        &self.party
    }
}
#[async_trait]
impl MPCNet for LeaderGpuNet {
    fn n_parties(&self) -> usize {
        self.party.n_parties
    }
    fn party_id(&self) -> u32 {
        0
    }
    fn is_init(&self) -> bool {
        true
    }
    fn get_comm(&self) -> (usize, usize) {
        self.party.get_comm()
    }
    fn add_comm(&self, _up: usize, _down: usize) {} // the synthetic counters live in libscz (LeaderSimNet, csrc/ctx.cu)
    async fn recv_from(&self, _id: u32, _sid: MultiplexedStreamID) -> Result<Bytes, MPCNetError> {
        Err(MPCNetError::NotConnected) // there are no peers in a `leader` build
    }
    async fn send_to(&self, _id: u32, _bytes: Bytes, _sid: MultiplexedStreamID) -> Result<(), MPCNetError> {
        Err(MPCNetError::NotConnected)
    }
}

// ------------------------------------------------------------------------------------------ NCCL over NVLink
/// One party per process and GPU (`hyperplonk/examples/bench_hyperplonk.rs:32-39` starts one process per party).
/// `uid` = NCCL's 128-byte unique id: `NcclGpuNet::unique_id()` on party 0, handed to the others by the host's own
/// channel -- e.g. the address file the reference already distributes (`--file`, mpc-net/src/multi.rs:109-140).
pub struct NcclGpuNet {
    party: GpuParty,
}
impl NcclGpuNet {
    pub fn unique_id() -> Result<[u8; SCZ_NCCL_UID_BYTES], MPCNetError> {
        let mut id = [0u8; SCZ_NCCL_UID_BYTES];
        match unsafe { scz_nccl_unique_id(id.as_mut_ptr() as *mut c_void) } {
            SCZ_OK => Ok(id),
            rc => Err(MPCNetError::Generic(format!("scz_nccl_unique_id failed ({rc})"))),
        }
    }
    pub fn new(device: i32, party_id: u32, n_parties: usize, uid: &[u8; SCZ_NCCL_UID_BYTES]) -> Result<Self, MPCNetError> {
        let mut ctx: *mut SczCtx = ptr::null_mut();
        let rc = unsafe { scz_ctx_create_nccl(device, party_id, n_parties as u32, uid.as_ptr() as *const c_void, &mut ctx) };
        if rc != SCZ_OK {
            return Err(MPCNetError::Generic(format!("scz_ctx_create_nccl failed ({rc})")));
        }
        Ok(Self { party: GpuParty { ctx, pp: Mutex::new(None), party_id, n_parties, lock: Mutex::new(()) } })
    }
}
impl GpuNet for NcclGpuNet {
    fn gpu(&self) -> &GpuParty {
        &self.party
    }
}
#[async_trait]
impl MPCNet for NcclGpuNet {
    fn n_parties(&self) -> usize {
        self.party.n_parties
    }
    fn party_id(&self) -> u32 {
        self.party.party_id
    }
    fn is_init(&self) -> bool {
        true
    }
    fn get_comm(&self) -> (usize, usize) {
        self.party.get_comm()
    }
    fn add_comm(&self, _up: usize, _down: usize) {}
    /// length-delimited like the reference's frames (u32 BE length + payload, multi.rs:29-35; here u64 LE): two
    /// ncclRecv on the ctx stream, staged through device memory
    async fn recv_from(&self, id: u32, _sid: MultiplexedStreamID) -> Result<Bytes, MPCNetError> {
        let p = &self.party;
        let _g = p.lock();
        let hdr = p.alloc(8)?;
        crate::check(p, unsafe { scz_net_recv(p.ctx(), id, hdr.ptr, 8) })?;
        let len = p.download::<u64>(&hdr, 1)?[0] as usize; // synchronises the stream
        let body = p.alloc(len)?;
        crate::check(p, unsafe { scz_net_recv(p.ctx(), id, body.ptr, len) })?;
        Ok(Bytes::from(p.download::<u8>(&body, len)?))
    }
    async fn send_to(&self, id: u32, bytes: Bytes, _sid: MultiplexedStreamID) -> Result<(), MPCNetError> {
        let p = &self.party;
        let _g = p.lock();
        let hdr = p.upload(&[bytes.len() as u64])?;
        crate::check(p, unsafe { scz_net_send(p.ctx(), id, hdr.ptr, 8) })?;
        let body = p.upload(&bytes[..])?;
        crate::check(p, unsafe { scz_net_send(p.ctx(), id, body.ptr, bytes.len()) })?;
        crate::check(p, unsafe { scz_ctx_sync(p.ctx()) }) // the buffers die with this scope
    }
}
