// Points rustc at the prebuilt libsbn254.so (python -m spartan_bn254_b200.build).
fn main() {
    let dir = std::env::var("SBN254_LIB_DIR").unwrap_or_else(|_| "../../spartan_bn254_b200/_lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sbn254");
    println!("cargo:rerun-if-env-changed=SBN254_LIB_DIR");
}
