// Links libkornia_b200.so.  KORNIA_B200_LIB_DIR points at the directory holding the shared library
// (kornia-rs_b200/lib in this repository).
fn main() {
    if let Ok(dir) = std::env::var("KORNIA_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=kornia_b200");
    println!("cargo:rerun-if-env-changed=KORNIA_B200_LIB_DIR");
}
