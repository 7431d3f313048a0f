// Links libscz.so (built in-tree by `make -C scalable-collaborative-zksnark_b200`).  SCZ_LIB_DIR overrides the search path.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("SCZ_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../scalable-collaborative-zksnark_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=scz");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=SCZ_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/scz.h");
}
