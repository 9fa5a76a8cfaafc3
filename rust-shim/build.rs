// Links libsar_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    let dir = std::env::var("SAR_B200_LIB_DIR")
        .unwrap_or_else(|_| "../strange-attractor-renderer_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sar_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=SAR_B200_LIB_DIR");
}
