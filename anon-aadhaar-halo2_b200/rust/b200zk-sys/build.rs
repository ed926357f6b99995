// Either link a prebuilt libb200zk.so (B200ZK_LIB_DIR) or compile the CUDA sources with nvcc
// through the `cc` crate.  No bindgen: the ABI is ~75 functions, declared by hand in lib.rs.
use std::{env, path::PathBuf};

fn main() {
    if let Ok(dir) = env::var("B200ZK_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=b200zk");
        return;
    }
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("csrc");
    cc::Build::new()
        .cuda(true)
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
        .flag("-O3").flag("-lineinfo").flag("-std=c++17").flag("--expt-relaxed-constexpr")
        .files(["core.cu", "ntt.cu", "msm.cu", "quotient.cu", "poly.cu", "permute.cu", "encoding.cu"].iter().map(|f| csrc.join(f)))
        .compile("b200zk");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
