//! Per-thread verification context.  The reference's types are Send + Sync and its functions re-entrant (no globals but two
//! read-only generators, M/amcl_utils.rs:26-30); the C ABI is re-entrant ACROSS contexts and a context is single-threaded, so
//! every thread that verifies gets its own b3_ctx, created on first use and destroyed with the thread.
use crate::b200::ffi::*;
use std::cell::RefCell;
use std::sync::atomic::{AtomicI32, Ordering};

static DEVICE: AtomicI32 = AtomicI32::new(0);

/// CUDA device used by contexts created AFTER this call (default 0; one process per GPU: set it to the local rank).
pub fn set_device(device: i32) {
    DEVICE.store(device, Ordering::Relaxed);
}

struct Holder(*mut b3_ctx);
impl Drop for Holder {
    fn drop(&mut self) {
        if !self.0.is_null() {
            unsafe { b3_ctx_destroy(self.0) }
        }
    }
}
thread_local! {
    static CTX: RefCell<Holder> = RefCell::new(Holder(std::ptr::null_mut()));
}

/// This thread's context, or None when no sm_100 device is usable: every verify then returns `false` -- the library has no
/// CPU fallback, and a verification failure is the only outcome the reference's bool-returning API can express.
pub(crate) fn with_ctx<T>(f: impl FnOnce(*mut b3_ctx) -> T) -> Option<T> {
    CTX.with(|c| {
        let mut h = c.borrow_mut();
        if h.0.is_null() {
            let mut p: *mut b3_ctx = std::ptr::null_mut();
            if unsafe { b3_ctx_create(DEVICE.load(Ordering::Relaxed), &mut p) } != B3_OK {
                return None;
            }
            h.0 = p;
        }
        Some(f(h.0))
    })
}

/// Device-resident table of decoded public keys (include/milagro_bls_b200.h: b3_keytable_*).  `PublicKey::from_bytes`
/// (decompression + key_validate, M/keys.rs:140-147) is paid once per validator; batch verification then names keys by
/// index.  Shared by every context of its device; read-only during verification.
pub struct KeyTable(pub(crate) *mut b3_keytable);
unsafe impl Send for KeyTable {}
unsafe impl Sync for KeyTable {}

impl KeyTable {
    pub fn with_capacity(capacity: usize) -> Option<KeyTable> {
        with_ctx(|ctx| {
            let mut t: *mut b3_keytable = std::ptr::null_mut();
            if unsafe { b3_keytable_create(ctx, capacity, &mut t) } == B3_OK { Some(KeyTable(t)) } else { None }
        })
        .flatten()
    }
    pub fn len(&self) -> usize {
        unsafe { b3_keytable_size(self.0) }
    }
    /// Appends 48-byte compressed keys with the checks of `PublicKey::from_bytes`; returns the index of the first key and one
    /// status per key (0 or the AmclError code; a rejected key keeps its slot and fails every set that names it).
    pub fn append_compressed(&mut self, keys48: &[u8]) -> Option<(usize, Vec<i32>)> {
        let n = keys48.len() / 48;
        let mut status = vec![0i32; n.max(1)];
        let mut first = 0usize;
        let rc = with_ctx(|ctx| unsafe { b3_keytable_append(ctx, self.0, keys48.as_ptr(), n, 1, 1, status.as_mut_ptr(), &mut first) })?;
        if rc != B3_OK {
            return None;
        }
        status.truncate(n);
        Some((first, status))
    }
}
impl Drop for KeyTable {
    fn drop(&mut self) {
        unsafe { b3_keytable_destroy(self.0) }
    }
}
