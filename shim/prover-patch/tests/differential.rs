//! prover/tests/cuda_differential.rs — the test that turns "bit-exact against our restatement" into "bit-exact against the reference".
//! Run on a machine with cargo (nightly-2025-05-09, rust-toolchain.toml:2), the stwo git dependency and a B200:
//!     NB200_LIB_DIR=<repo>/nexus_zkvm_b200 cargo test -p nexus-vm-prover --features cuda --test cuda_differential -- --test-threads 1
use nexus_common_testing::program_trace;
use nexus_vm::trace::k_trace_direct;
use nexus_vm_prover::machine::{BaseComponent, Machine};

fn one(log_size: u32) {
    let (view, trace) = k_trace_direct(&program_trace(log_size), 1).unwrap();      // common-testing/src/lib.rs:5-30: the ADD chain
    let simd = Machine::<BaseComponent>::prove(&trace, &view).unwrap();            // Stwo SimdBackend (the reference)
    let cuda = Machine::<BaseComponent>::prove_cuda(&[], &trace, &view).unwrap();  // libnexus_b200
    // roots first: a mismatch in tree 0 isolates iFFT / LDE / Blake2s-Merkle (flip nb200_set_flavor(merkle_hash) if only the hash differs)
    assert_eq!(simd.stark_proof.commitments, cuda.stark_proof.commitments, "Merkle roots differ at log_size {log_size}");
    assert_eq!(simd.claimed_sum, cuda.claimed_sum, "LogUp claimed sums differ");
    assert_eq!(postcard::to_allocvec(&simd.stark_proof).unwrap(), postcard::to_allocvec(&cuda.stark_proof).unwrap(), "proof bytes differ");
    // and the reference verifier accepts the GPU proof (incl. its own recomputation of the tree-0 root, machine.rs:363-417)
    nexus_vm_prover::verify(cuda, &view).unwrap();
}

#[test] fn cuda_matches_simd_backend_small() { for l in 8..=12 { one(l) } }
#[test] fn cuda_matches_simd_backend_2_16() { one(16) }
#[test] #[ignore = "minutes on the CPU side"] fn cuda_matches_simd_backend_2_20() { one(20) }

/// golden vectors for the repository's oracle: run once, commit the output under tests/golden/ (tools/check_golden.py reads it)
#[test] #[ignore]
fn dump_golden_vectors() {
    for l in 8..=14 {
        let (view, trace) = k_trace_direct(&program_trace(l), 1).unwrap();
        let p = Machine::<BaseComponent>::prove(&trace, &view).unwrap();
        let roots: Vec<String> = p.stark_proof.commitments.iter().map(|h| h.0.iter().map(|b| format!("{b:02x}")).collect()).collect();
        println!("{{\"log_size\": {l}, \"roots\": {roots:?}, \"proof_postcard_hex\": \"{}\"}}",
                 postcard::to_allocvec(&p.stark_proof).unwrap().iter().map(|b| format!("{b:02x}")).collect::<String>());
    }
}
