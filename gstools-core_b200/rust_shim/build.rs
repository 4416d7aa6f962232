// build.rs -- compile the hand-written CUDA kernels for sm_100a and link them into the crate.
// UNBUILT in this repository (no Rust toolchain in the image); see README.md.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("GSFIELD_CSRC").unwrap_or_else(|_| "csrc".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("libgsfield.a");
    let obj = out.join("gsfield.o");

    let status = Command::new(&nvcc)
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
        .args(["-Xcompiler", "-fPIC", "-c"])
        .arg(csrc.join("gsfield.cu"))
        .arg("-o")
        .arg(&obj)
        .status()
        .expect("nvcc not found: the CUDA path has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    let status = Command::new("ar").arg("crs").arg(&lib).arg(&obj).status().unwrap();
    assert!(status.success());

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=gsfield");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
    // gsfield.cu includes every .cuh / .inc next to it: watch the whole directory
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-env-changed=NVCC");
    println!("cargo:rerun-if-env-changed=GSFIELD_CSRC");
}
