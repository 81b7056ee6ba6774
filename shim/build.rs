fn main() {
    if let Ok(dir) = std::env::var("EE_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=ee_b200");
}
