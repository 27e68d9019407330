//! amcl points <-> the uncompressed ZCash encodings the C ABI speaks (A/bls381/core.rs:177-190, 344-364: the reference's own
//! serialize_uncompressed_g1 / _g2 and deserialize_g1 / _g2).
use crate::amcl_utils::{deserialize_g1, deserialize_g2, AmclError, GroupG1, GroupG2, G1_BYTES, G2_BYTES};
use crate::b200::ffi::*;
use crate::BLSCurve::bls381::utils::{serialize_uncompressed_g1, serialize_uncompressed_g2};

pub(crate) const G1_WIRE: usize = 2 * G1_BYTES; // 96
pub(crate) const G2_WIRE: usize = 2 * G2_BYTES; // 192

pub(crate) fn g1_wire(p: &GroupG1) -> [u8; G1_WIRE] {
    serialize_uncompressed_g1(p)
}
pub(crate) fn g2_wire(p: &GroupG2) -> [u8; G2_WIRE] {
    serialize_uncompressed_g2(p)
}
pub(crate) fn g1_from_wire(b: &[u8; G1_WIRE]) -> Result<GroupG1, AmclError> {
    deserialize_g1(b)
}
pub(crate) fn g2_from_wire(b: &[u8; G2_WIRE]) -> Result<GroupG2, AmclError> {
    deserialize_g2(b)
}
/// status code of the C ABI -> the reference's error enum (A/errors.rs:1-11, same order as the B3_ERR_* constants)
pub(crate) fn amcl_error(code: i32) -> AmclError {
    match code {
        B3_ERR_AGGREGATE_EMPTY_POINTS => AmclError::AggregateEmptyPoints,
        B3_ERR_INVALID_G1_SIZE => AmclError::InvalidG1Size,
        B3_ERR_INVALID_G2_SIZE => AmclError::InvalidG2Size,
        B3_ERR_INVALID_YFLAG => AmclError::InvalidYFlag,
        _ => AmclError::InvalidPoint,
    }
}
