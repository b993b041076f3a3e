//! Link against libnexus_b200.so.  NB200_LIB_DIR points at the directory that holds it
//! (`python -m nexus_zkvm_b200.build` writes nexus_zkvm_b200/libnexus_b200.so); the CUDA runtime is linked statically
//! into the library, so nothing else is needed at link time.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("NB200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        // default: the in-tree build of this repository (shim/nexus-b200 -> ../../nexus_zkvm_b200)
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../nexus_zkvm_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=nexus_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=NB200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/nb200.h");
}
