// libclq.so is built by `make -C clique_b200/csrc` (nvcc -gencode arch=compute_100a,code=sm_100a); point CLQ_LIB_DIR at the
// directory that holds it (clique_b200/).
fn main() {
    let dir = std::env::var("CLQ_LIB_DIR").expect("set CLQ_LIB_DIR to the directory that holds libclq.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=clq");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=CLQ_LIB_DIR");
}
