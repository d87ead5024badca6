// Links libazb200.so (built by `python alphazero-rs_b200/build.py` with nvcc for sm_100a).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("AZB200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../alphazero-rs_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=azb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=AZB200_LIB_DIR");
}
