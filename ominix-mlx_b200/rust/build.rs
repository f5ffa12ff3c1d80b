//! Builds the CUDA kernels with nvcc for sm_100a and links them statically.
//! Counterpart of mlx-rs/mlx-sys/build.rs (which drives cmake + bindgen over mlx-c); here the C ABI
//! is small enough to be declared by hand in src/ffi.rs, so no bindgen step is needed.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let sources = ["omx_api.cu", "rope.cu", "norm.cu", "prologue.cu", "kv_cache.cu", "sdpa_generic.cu", "decode.cu", "fmha_sm100.cu"];
    let mut objects = Vec::new();
    for src in sources {
        let obj = out.join(src.replace(".cu", ".o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(src))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("failed to run nvcc (sm_100a toolchain required; there is no CPU fallback)");
        assert!(status.success(), "nvcc failed on {src}");
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
        objects.push(obj);
    }
    let lib = out.join("libomx_attn.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objects).status().expect("ar");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=omx_attn");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
    println!("cargo:rerun-if-changed={}", manifest.join("../../include/omx_attn.h").display());
}
