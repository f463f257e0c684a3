// build.rs — compiles the hand-written sm_100a kernels with nvcc and links them into the crate.
// No CPU fallback is built: without nvcc the build fails loudly.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    let nvcc = format!("{cuda}/bin/nvcc");
    // the same list as SRCS in basic_sparse_matrix_b200/csrc/Makefile (tests/test_rust_ffi.py compares the two)
    let sources = ["runtime.cu", "handles.cu", "planner.cu", "dispatch.cu", "pipeline.cu", "solve.cu", "gen_api.cu", "spmm_rows.cu",
                   "spmm_rows_f64.cu", "spmm_rows_f32.cu", "spmm_merge.cu", "spmm_rowblock.cu", "convert.cu", "gen.cu", "bsm_nccl.cu"];
    let mut objects = Vec::new();
    for src in sources {
        let obj = out.join(format!("{src}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "-Xfatbin", "-compress-all", "-Xcompiler", "-fPIC", "-Icsrc", "-c"])
            .arg(format!("csrc/{src}"))
            .arg("-o").arg(&obj)
            .status()
            .expect("nvcc not found: the gpu module has no CPU fallback");
        assert!(status.success(), "nvcc failed on {src}");
        println!("cargo:rerun-if-changed=csrc/{src}");
        objects.push(obj);
    }
    let lib = out.join("libbsm_b200.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objects).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=bsm_b200");
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=nccl");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    for hdr in ["bsm_common.cuh", "bsm_internal.h", "line_length.h", "kernels.h", "spmm_stream.cuh", "spmm_rows_kernel.cuh", "spmm_rows_inst.cuh", "bsm.h"] {
        println!("cargo:rerun-if-changed=csrc/{hdr}");
    }
}
