// Links libs2k_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"` in the B200 repository).
fn main() {
    let dir = std::env::var("S2K_LIB_DIR").expect("set S2K_LIB_DIR to the directory holding libs2k_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=s2k_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=S2K_LIB_DIR");
}
