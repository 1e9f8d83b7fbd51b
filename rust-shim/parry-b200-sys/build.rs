// Points the linker at libparry_b200.so (built by `python -m parry_b200.build`, in-tree: <repo>/parry_b200/).
fn main() {
    println!("cargo:rerun-if-env-changed=PARRY_B200_LIB_DIR");
    let dir = std::env::var("PARRY_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        manifest.join("../../parry_b200").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=parry_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
}
