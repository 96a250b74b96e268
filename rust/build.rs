// build.rs — UNVERIFIED (never compiled here: no cargo/rustc in the image).
// Compiles the CUDA sources for sm_100a with nvcc and links them (plus the CUDA runtime) into the crate.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from("friendship-b200/libfriendship_b200/csrc");
    // the same translation units as __graft_entry__.py (the C++ Dispatch restatement host/dispatch.cc is not needed:
    // Dispatch stays in Rust).  interp_device_src.cc is generated from interp_device.inc by __graft_entry__.py.
    let units = [("capi.cu", false), ("renderer.cu", false), ("multi.cu", false), ("interp.cu", false), ("flatten.cc", false),
                 ("jit.cc", false), ("interp_device_src.cc", false), ("osc.cu", true), ("scan.cu", true)];
    let mut objs = vec![];
    for (src, fmad) in units.iter() {
        let obj = out.join(format!("{}.o", src));
        let status = Command::new("nvcc")
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                    "-Xcompiler", "-fPIC", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
                    if *fmad { "-fmad=true" } else { "-fmad=false" }, "-x", "cu", "-c"])
            .arg(csrc.join(src)).arg("-o").arg(&obj)
            .status().expect("nvcc not found");
        assert!(status.success(), "nvcc failed on {}", src);
        objs.push(obj);
    }
    let lib = out.join("libfriendship_b200.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=friendship_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
    println!("cargo:rustc-link-lib=dl");          // the stage JIT dlopens NVRTC and the driver API
    println!("cargo:rerun-if-changed=friendship-b200/libfriendship_b200/csrc");
}
