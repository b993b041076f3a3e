//! `nexus-b200`: the Rust side of the drop-in boundary described in INTEGRATION.md.
//!
//! * [`ffi`]      — `extern "C"` declarations, one per entry point of `include/nb200.h` (checked against the header by
//!                  `tests/test_shim_ffi_cpu.py` in the repository: same names, same arity).
//! * [`context`]  — owning wrappers (`Context`, `Columns`, `Channel`, `Scheme`, `Air`) with `Result`-returning methods.
//! * [`recorder`] — `impl EvalAtRow for Recorder`: running `FrameworkEval::evaluate(recorder)` on the reference's
//!                  `MachineEval<C>` (prover/src/components/mod.rs:48-57) yields the SSA bytecode `nb200_air_load` takes.
//!                  `nexus_zkvm_b200/air.py` is the executable specification of the format and of the emission order.
//!
//! The crate was written in an image without `cargo`/`rustc`, against stwo @0790eba as restated in
//! SURVEY.md Appendix A: expect to touch trait-bound lists on the first `cargo check`.  `shim/prover-patch` holds the module
//! that plugs it into `nexus_vm_prover::Machine` and the differential test against `SimdBackend`.
pub mod context;
pub mod ffi;
pub mod recorder;

pub use context::{Air, Channel, Columns, Context, Error, PcsParams, Scheme};
pub use recorder::{AirBytecode, Recorder, RelationParams};
