// build.rs — links liblorikeet_b200.so (built by lorikeet_b200/csrc/build.sh with nvcc for sm_100a).
// LORIKEET_B200_LIB_DIR = the directory that holds the shared object (default: ../../lorikeet_b200 of this checkout).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("LORIKEET_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../lorikeet_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=lorikeet_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=LORIKEET_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/lorikeet_b200.h");
}
