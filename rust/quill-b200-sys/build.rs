// Links libquill_b200.so (built by `make -C quill_zkvm_b200/csrc`: nvcc -gencode arch=compute_100a,code=sm_100a).
// QUILL_B200_LIB_DIR overrides the directory; the default is the in-tree location relative to this crate.
fn main() {
    let dir = std::env::var("QUILL_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        here.join("../../quill_zkvm_b200").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=quill_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=QUILL_B200_LIB_DIR");
}
