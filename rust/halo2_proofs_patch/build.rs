// build.rs of the patched halo2_proofs fork (feature "gpu"): link libtaiga_b200.so and generate the bindings of
// include/taiga_b200.h.  Replaces nothing in the reference; the reference has no FFI on this path.
fn main() {
    if std::env::var("CARGO_FEATURE_GPU").is_err() {
        return;
    }
    let lib_dir = std::env::var("TAIGA_B200_LIB_DIR").expect("TAIGA_B200_LIB_DIR = directory holding libtaiga_b200.so");
    let inc_dir = std::env::var("TAIGA_B200_INCLUDE").expect("TAIGA_B200_INCLUDE = directory holding taiga_b200.h");
    println!("cargo:rustc-link-search=native={lib_dir}");
    println!("cargo:rustc-link-lib=dylib=taiga_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{lib_dir}");
    println!("cargo:rerun-if-changed={inc_dir}/taiga_b200.h");
    let bindings = bindgen::Builder::default()
        .header(format!("{inc_dir}/taiga_b200.h"))
        .allowlist_function("tb_.*")
        .allowlist_type("tb_.*")
        .allowlist_var("TB_.*")
        .generate()
        .expect("bindgen over taiga_b200.h");
    let out = std::path::PathBuf::from(std::env::var("OUT_DIR").unwrap());
    bindings.write_to_file(out.join("tb.rs")).expect("write bindings");
}
