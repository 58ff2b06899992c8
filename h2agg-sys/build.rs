// Finds libh2agg.so: H2AGG_LIB_DIR, else the in-tree build next to this crate (../halo2-aggregation_b200).
// The library itself is built by `python halo2-aggregation_b200/_build.py` (nvcc, sm_100a); this script only links it.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("H2AGG_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..").join("halo2-aggregation_b200")
    });
    if !dir.join("libh2agg.so").exists() {
        panic!("libh2agg.so not found in {} (set H2AGG_LIB_DIR or run python halo2-aggregation_b200/_build.py)", dir.display());
    }
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=h2agg");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=H2AGG_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/h2agg.h");
}
