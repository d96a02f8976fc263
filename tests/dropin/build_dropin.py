"""Builds tests/dropin/_build/gsdf_dropin.so: the reference's UNMODIFIED src/gaussian_rasterizer.cpp (autograd glue,
SURVEY.md row a-19) + tests/dropin/dropin_module.cpp, linked against this repository's libgsr_torch_shim.so.

The reference sources are compiled where they lie: a throw-away directory of SYMLINKS (`_build/refview`) gives the
compiler the reference's include/gaussian_rasterizer.h and include/rasterize_points.h next to the stub
tests/dropin/stub/gaussian_model.h (the real header needs Eigen / Sophus / OpenCV, absent here); nothing is copied and
the view is removed after the build.  No-op when /root/reference is absent (the GPU box uses the prebuilt file)."""
import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GSDF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_build")
SO = os.path.join(OUT, "gsdf_dropin.so")
STAMP = os.path.join(OUT, ".stamp")
PKG = os.path.join(ROOT, "gsdf_slam_b200")


def main() -> int:
    src = os.path.join(REF, "src", "gaussian_rasterizer.cpp")
    if not os.path.exists(src):
        print(f"build_dropin: {REF} not present; keeping prebuilt {SO} (if any)")
        return 0
    shim = os.path.join(PKG, "libgsr_torch_shim.so")
    if not os.path.exists(shim):
        print("build_dropin: libgsr_torch_shim.so missing (build the shim first)")
        return 1
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    for f in (src, os.path.join(REF, "include", "gaussian_rasterizer.h"), os.path.join(REF, "include", "rasterize_points.h"),
              os.path.join(HERE, "dropin_module.cpp"), os.path.join(HERE, "stub", "gaussian_model.h"), os.path.abspath(__file__)):
        h.update(open(f, "rb").read())
    stamp = h.hexdigest()
    if os.path.exists(SO) and os.path.exists(STAMP) and open(STAMP).read().strip() == stamp:
        print("build_dropin: up to date")
        return 0
    view = os.path.join(OUT, "refview")
    shutil.rmtree(view, ignore_errors=True)
    os.makedirs(os.path.join(view, "include"))
    os.makedirs(os.path.join(view, "src"))
    for name in ("gaussian_rasterizer.h", "rasterize_points.h"):
        os.symlink(os.path.join(REF, "include", name), os.path.join(view, "include", name))
    os.symlink(os.path.join(HERE, "stub", "gaussian_model.h"), os.path.join(view, "include", "gaussian_model.h"))
    os.symlink(src, os.path.join(view, "src", "gaussian_rasterizer.cpp"))
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{view}"] + [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}",
                                                                   "-I/usr/local/cuda/include"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-std=c++17", "-O2", abi, "-DTORCH_EXTENSION_NAME=gsdf_dropin", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-fPIC", "-w", "-shared", *inc, os.path.join(view, "src", "gaussian_rasterizer.cpp"),
           os.path.join(HERE, "dropin_module.cpp"), "-o", SO, f"-L{PKG}", "-l:libgsr_torch_shim.so", "-l:libgsr.so",
           f"-L{libdir}", "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-Wl,--no-as-needed",
           "-Wl,-rpath,$ORIGIN/../../../gsdf_slam_b200", f"-Wl,-rpath,{libdir}"]
    rc = subprocess.call(cmd)
    shutil.rmtree(view, ignore_errors=True)
    if rc:
        return rc
    open(STAMP, "w").write(stamp)
    print(f"build_dropin: built {SO}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
