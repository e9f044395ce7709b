// links libpathplanning_b200.so; PATHPLANNING_B200_LIB_DIR points at the directory that holds it
fn main() {
    if let Ok(dir) = std::env::var("PATHPLANNING_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=pathplanning_b200");
}
