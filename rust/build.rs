// bellman/build.rs -- links libbellman_b200.so (sm_100a) when the `cuda` feature is enabled.
// BELLMAN_B200_DIR = path of the bellman-b200 checkout (the directory holding `Makefile`).
use std::env;
use std::process::Command;

fn main() {
    if env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let dir = env::var("BELLMAN_B200_DIR").expect("BELLMAN_B200_DIR: path of the bellman-b200 checkout");
    // nvcc -gencode arch=compute_100a,code=sm_100a ... -> bellman_mpc_b200/libbellman_b200.so
    let ok = Command::new("make")
        .args(&["-j8", "-C", &dir])
        .status()
        .expect("running make")
        .success();
    assert!(ok, "building libbellman_b200.so failed");
    println!("cargo:rustc-link-search=native={}/bellman_mpc_b200", dir);
    println!("cargo:rustc-link-lib=dylib=bellman_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}/bellman_mpc_b200", dir);
    println!("cargo:rerun-if-changed={}/include/bellman_b200.h", dir);
    println!("cargo:rerun-if-env-changed=BELLMAN_B200_DIR");
}
