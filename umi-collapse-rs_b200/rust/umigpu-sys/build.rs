// UMIGPU_LIB_DIR = directory holding libumigpu.so (umi-collapse-rs_b200/csrc after `make`)
fn main() {
    let dir = std::env::var("UMIGPU_LIB_DIR").unwrap_or_else(|_| "../../csrc".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=umigpu");
    println!("cargo:rerun-if-env-changed=UMIGPU_LIB_DIR");
}
