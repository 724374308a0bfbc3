// build.rs — compiles the CUDA sources of the C ABI (include/rustradio_cuda.h) with nvcc for
// sm_100a and links them into the crate.  UNCOMPILED in this repository's environment (no cargo).
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("../../..");
    let csrc = root.join("rustradio_b200/csrc");
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let nvcc = std::env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    // keep in step with SRCS in rustradio_b200/csrc/Makefile (tests/test_abi.py::test_rust_build_lists_every_source)
    let srcs = ["runtime.cu", "fir.cu", "fir_tc.cu", "fir_tc5.cu", "fir_tcc.cu", "fftfilt.cu", "fftfilt_fold.cu", "fftfilt_poly.cu", "resample.cu", "ingest.cu",
                "fftstream.cu", "hilbert.cu", "elementwise.cu", "blocks.cu", "blocks_capi.cu", "sources.cu"];
    let mut objs = vec![];
    for s in srcs {
        let o = out.join(format!("{s}.o"));
        let st = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(s)).arg("-o").arg(&o)
            .status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {s}");
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
        objs.push(o);
    }
    let lib = out.join("librustradio_cuda.a");
    let st = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap();
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=rustradio_cuda");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
}
