// Links libb2bu.so (built by `python -m basisu_rs_b200.build`, it sits in basisu_rs_b200/).
// B2BU_LIB_DIR overrides the search directory.  Source only: not compiled in this image (no rustc).
fn main() {
    let dir = std::env::var("B2BU_LIB_DIR").unwrap_or_else(|_| {
        let here = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        here.join("..").join("basisu_rs_b200").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b2bu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=B2BU_LIB_DIR");
}
