//! Replacement bodies for `impl PublicKey` in the reference's src/keys.rs (lines 140-186).  `SecretKey`, `Keypair` and
//! `PublicKey::from_secret_key` (keys.rs:24-137, 189-215) are untouched: secret material never goes to the GPU.
use crate::amcl_utils::{AmclError, G1_BYTES};
use crate::b200::ctx::with_ctx;
use crate::b200::ffi::*;
use crate::b200::wire::*;
use crate::keys::PublicKey;

impl PublicKey {
    /// keys.rs:140-147: decompression, not-infinity and the G1 subgroup test in one kernel
    pub fn from_bytes(bytes: &[u8]) -> Result<PublicKey, AmclError> {
        Self::decode(bytes, 1)
    }
    /// keys.rs:150-155
    pub fn from_bytes_unchecked(bytes: &[u8]) -> Result<PublicKey, AmclError> {
        Self::decode(bytes, 0)
    }
    fn decode(bytes: &[u8], validate: i32) -> Result<PublicKey, AmclError> {
        if bytes.len() != G1_BYTES {
            return Err(AmclError::InvalidG1Size);
        }
        let mut out = [0u8; G1_WIRE];
        let mut st = 0i32;
        let rc = with_ctx(|ctx| unsafe { b3_g1_decompress(ctx, bytes.as_ptr(), 1, validate, out.as_mut_ptr(), &mut st) });
        match rc {
            Some(B3_OK) if st == 0 => Ok(PublicKey { point: g1_from_wire(&out)? }),
            Some(B3_OK) => Err(amcl_error(st)),
            _ => Err(AmclError::InvalidPoint),
        }
    }
    /// Batched form for a validator set: one launch for all keys, one Result per key.
    pub fn from_bytes_batch(keys48: &[u8]) -> Vec<Result<PublicKey, AmclError>> {
        let n = keys48.len() / G1_BYTES;
        let mut out = vec![0u8; G1_WIRE * n.max(1)];
        let mut st = vec![0i32; n.max(1)];
        let rc = with_ctx(|ctx| unsafe { b3_g1_decompress(ctx, keys48.as_ptr(), n, 1, out.as_mut_ptr(), st.as_mut_ptr()) });
        (0..n)
            .map(|i| {
                if rc != Some(B3_OK) {
                    return Err(AmclError::InvalidPoint);
                }
                if st[i] != 0 {
                    return Err(amcl_error(st[i]));
                }
                let mut w = [0u8; G1_WIRE];
                w.copy_from_slice(&out[G1_WIRE * i..G1_WIRE * (i + 1)]);
                Ok(PublicKey { point: g1_from_wire(&w)? })
            })
            .collect()
    }
    /// keys.rs:158-160
    pub fn as_bytes(&self) -> [u8; G1_BYTES] {
        let w = g1_wire(&self.point);
        let mut out = [0u8; G1_BYTES];
        let mut st = 0i32;
        let rc = with_ctx(|ctx| unsafe { b3_g1_compress(ctx, w.as_ptr(), 1, out.as_mut_ptr(), &mut st) });
        debug_assert!(rc == Some(B3_OK) && st == 0);
        out
    }
    /// keys.rs:181-186
    pub fn key_validate(&self) -> bool {
        let w = g1_wire(&self.point);
        let (mut st, mut valid) = (0i32, 0i32);
        let rc = with_ctx(|ctx| unsafe { b3_g1_validate(ctx, w.as_ptr(), 1, &mut st, &mut valid) });
        rc == Some(B3_OK) && st == 0 && valid == 1
    }
}
