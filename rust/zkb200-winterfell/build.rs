// Links libzkb200.so; set ZKB200_LIB_DIR to the directory that holds it (zk_stark_project_b200/ in this repo).
fn main() {
    let dir = std::env::var("ZKB200_LIB_DIR").unwrap_or_else(|_| "../../zk_stark_project_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zkb200");
    println!("cargo:rerun-if-env-changed=ZKB200_LIB_DIR");
}
