// NOT COMPILED OR RUN: this image has no cargo/rustc.  Written against the reference's sources (bvpav/raydar) and
// include/raydar_cuda.h; the same C ABI is exercised through ctypes by tests/ and bench.py.  See INTEGRATION.md.
fn main() {
    // directory that holds libraydar_cuda.so (python -m raydar_b200.build puts it in raydar_b200/)
    let dir = std::env::var("RAYDAR_CUDA_LIB_DIR").expect("set RAYDAR_CUDA_LIB_DIR");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=raydar_cuda");
    println!("cargo:rerun-if-env-changed=RAYDAR_CUDA_LIB_DIR");
}
