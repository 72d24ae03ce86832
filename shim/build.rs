fn main() {
    // RSRL_B200_LIB_DIR = <repo>/rsrl_b200/csrc
    let dir = std::env::var("RSRL_B200_LIB_DIR").expect("set RSRL_B200_LIB_DIR to the directory of librsrl_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rsrl_b200");
}
