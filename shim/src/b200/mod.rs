//! The B200 back end of the verification path: `mod b200;` in the reference's src/lib.rs (next to `mod aggregates;`,
//! M/lib.rs:12-15) plus `pub use b200::{set_device, KeyTable};`.
pub mod ctx;
pub mod ffi;
pub mod wire;
pub use ctx::{set_device, KeyTable};
