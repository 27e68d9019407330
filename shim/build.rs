// Links the CUDA library built by `python -c 'import __graft_entry__ as g; g.build()'` (milagro_bls_b200/libmilagro_bls_b200.so).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("MILAGRO_BLS_B200_LIB_DIR")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../milagro_bls_b200"));
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=milagro_bls_b200");
    println!("cargo:rerun-if-env-changed=MILAGRO_BLS_B200_LIB_DIR");
}
