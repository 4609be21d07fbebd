// Builds libbamgpu.so with nvcc for sm_100a (the repo's Makefile) and links it.
// UNBUILT in this repository's image (no cargo/rustc there).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let status = Command::new("make")
        .arg("-C").arg(&root).arg("bam_parallel_b200/lib/libbamgpu.so")
        .status().expect("make (nvcc) failed to start");
    assert!(status.success(), "building libbamgpu.so failed: there is no CPU fallback");
    let lib = root.join("bam_parallel_b200/lib");
    println!("cargo:rustc-link-search=native={}", lib.display());
    println!("cargo:rustc-link-lib=dylib=bamgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib.display());
    println!("cargo:rerun-if-changed={}", root.join("include/bamgpu.h").display());
    println!("cargo:rerun-if-changed={}", root.join("bam_parallel_b200/csrc").display());
}
