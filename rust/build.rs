// Links the in-tree CUDA library built by `make` at the repository root.
fn main() {
    let dir = std::env::var("BEVYRAY_B200_LIB_DIR").unwrap_or_else(|_| "../bevyray_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=bevyray_b200");
    println!("cargo:rerun-if-env-changed=BEVYRAY_B200_LIB_DIR");
}
