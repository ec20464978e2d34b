// Links the prebuilt CUDA library.  Point VPBS_COMMIT_LIB_DIR at the directory that holds
// libvpbs_commit.so (verifiable-fhe-paper_b200/ in this repository).
fn main() {
    let dir = std::env::var("VPBS_COMMIT_LIB_DIR")
        .expect("set VPBS_COMMIT_LIB_DIR to the directory containing libvpbs_commit.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=vpbs_commit");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=VPBS_COMMIT_LIB_DIR");
}
