// Links libpanopaea_b200 (default: the shared library built by `python -m panopaea_b200.build`), or, with the
// `build-cuda` feature, compiles the .cu sources with nvcc through the `cc` crate -- "cc-built .cu objects" as
// BASELINE.json's north star puts it.  The flags are the ones panopaea_b200/build.py uses; --fmad=false matters:
// rustc never contracts a*b+c, and the kernels reproduce the reference's element-wise results bit for bit.
use std::env;
use std::path::PathBuf;

fn repo_root() -> PathBuf {
    match env::var("PANOPAEA_B200_ROOT") {
        Ok(p) => PathBuf::from(p),
        // rust/panopaea-b200-sys -> repository root
        Err(_) => PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..").join(".."),
    }
}

#[cfg(feature = "build-cuda")]
fn build_or_link(root: &PathBuf) {
    let csrc = root.join("panopaea_b200").join("csrc");
    let mut b = cc::Build::new();
    b.cuda(true)
        .cpp(true)
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-std=c++17")
        .flag("-O3")
        .flag("-lineinfo")
        .flag("--fmad=false")
        .include(root.join("include"));
    for entry in std::fs::read_dir(&csrc).expect("panopaea_b200/csrc") {
        let p = entry.unwrap().path();
        if p.extension().map_or(false, |e| e == "cu") {
            println!("cargo:rerun-if-changed={}", p.display());
            b.file(p);
        }
    }
    b.compile("panopaea_b200");                       // static archive, linked into the crate
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}

#[cfg(not(feature = "build-cuda"))]
fn build_or_link(root: &PathBuf) {
    let lib = root.join("panopaea_b200").join("lib");
    println!("cargo:rustc-link-search=native={}", lib.display());
    println!("cargo:rustc-link-lib=dylib=panopaea_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib.display());
}

fn main() {
    let root = repo_root();
    println!("cargo:rerun-if-env-changed=PANOPAEA_B200_ROOT");
    println!("cargo:rerun-if-changed={}", root.join("include").join("panopaea_b200.h").display());
    build_or_link(&root);
}
