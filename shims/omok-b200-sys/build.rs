// Links libomok_b200.so (built by `make -C omok-ai_b200/csrc`); OMOK_B200_LIB_DIR = the directory that holds it.
fn main() {
    let dir = std::env::var("OMOK_B200_LIB_DIR").expect("set OMOK_B200_LIB_DIR to the directory of libomok_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=omok_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=OMOK_B200_LIB_DIR");
}
