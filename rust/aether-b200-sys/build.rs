fn main() {
    let dir = std::env::var("AETHER_B200_LIB_DIR").expect("set AETHER_B200_LIB_DIR to the directory holding libaether_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=aether_b200"); // the CUDA runtime is linked statically inside
}
