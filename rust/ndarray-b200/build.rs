// Builds libndb200.so with nvcc for sm_100a (same flags as ndarray_b200/build.py) unless NDB200_LIB_DIR
// points at a prebuilt one, then links it.  No other backend exists: without nvcc or a prebuilt library
// the build fails.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let lib_dir = match env::var("NDB200_LIB_DIR") {
        Ok(d) => PathBuf::from(d),
        Err(_) => {
            let out = PathBuf::from(env::var("OUT_DIR").unwrap());
            let csrc = root.join("ndarray_b200/csrc");
            let mut objs = vec![];
            for entry in std::fs::read_dir(&csrc).unwrap() {
                let p = entry.unwrap().path();
                if p.extension().map(|e| e == "cu").unwrap_or(false) {
                    let o = out.join(p.file_stem().unwrap()).with_extension("o");
                    let st = Command::new("nvcc")
                        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                               "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-Xcompiler", "-fPIC",
                               "--expt-relaxed-constexpr", "-c"])
                        .arg(&p).arg("-o").arg(&o).status().expect("nvcc is required: there is no CPU build");
                    assert!(st.success(), "nvcc failed on {:?}", p);
                    println!("cargo:rerun-if-changed={}", p.display());
                    objs.push(o);
                }
            }
            let so = out.join("libndb200.so");
            let st = Command::new("nvcc").args(["-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o"])
                .arg(&so).args(&objs).status().unwrap();
            assert!(st.success());
            out
        }
    };
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=ndb200");
}
