//! Replacement bodies for `impl Signature` in the reference's src/signature.rs (lines 27-51).  `Signature::new`
//! (signature.rs:17-24, signing) is untouched.
use crate::amcl_utils::{AmclError, G2_BYTES};
use crate::b200::ctx::with_ctx;
use crate::b200::ffi::*;
use crate::b200::wire::*;
use crate::keys::PublicKey;
use crate::signature::Signature;

impl Signature {
    /// signature.rs:27-40: subgroup check of the signature, key_validate-free pairing check e(sig, -G1) e(H(msg), pk) == 1
    pub fn verify(&self, msg: &[u8], pk: &PublicKey) -> bool {
        let (s, k) = (g2_wire(&self.point), g1_wire(&pk.point));
        let mut accept = 0i32;
        let rc = with_ctx(|ctx| unsafe { b3_verify(ctx, s.as_ptr(), k.as_ptr(), msg.as_ptr(), msg.len(), &mut accept, std::ptr::null_mut()) });
        rc == Some(B3_OK) && accept == 1
    }
    /// signature.rs:43-46 (no subgroup check here: `verify` does it)
    pub fn from_bytes(bytes: &[u8]) -> Result<Signature, AmclError> {
        if bytes.len() != G2_BYTES {
            return Err(AmclError::InvalidG2Size);
        }
        let mut out = [0u8; G2_WIRE];
        let mut st = 0i32;
        let rc = with_ctx(|ctx| unsafe { b3_g2_decompress(ctx, bytes.as_ptr(), 1, out.as_mut_ptr(), &mut st) });
        match rc {
            Some(B3_OK) if st == 0 => Ok(Signature { point: g2_from_wire(&out)? }),
            Some(B3_OK) => Err(amcl_error(st)),
            _ => Err(AmclError::InvalidPoint),
        }
    }
    /// signature.rs:49-51
    pub fn as_bytes(&self) -> [u8; G2_BYTES] {
        let w = g2_wire(&self.point);
        let mut out = [0u8; G2_BYTES];
        let mut st = 0i32;
        let rc = with_ctx(|ctx| unsafe { b3_g2_compress(ctx, w.as_ptr(), 1, out.as_mut_ptr(), &mut st) });
        debug_assert!(rc == Some(B3_OK) && st == 0);
        out
    }
}
