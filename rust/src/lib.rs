//! `src/cuda/mod.rs` of the reference once this directory is dropped in: the B200 prover behind `machine.prove::<P>`.
pub mod ffi;
pub mod flatten;
pub mod proof;
pub mod prover;
