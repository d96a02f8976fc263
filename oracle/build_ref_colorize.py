"""TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/gsdf_ref_colorize.so: the reference's UNMODIFIED
cuda_colorize/colorize_pointcloud.cu (compiled where it lies under /root/reference, never copied) plus the
pybind11 wrapper oracle/ref_colorize.cpp, against this image's LibTorch.  No-op when /root/reference is
absent (the GPU box uses the prebuilt file) or when the output is up to date."""
import hashlib
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GSDF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SRC = os.path.join(REF, "cuda_colorize", "colorize_pointcloud.cu")
WRAP = os.path.join(HERE, "ref_colorize.cpp")
SO = os.path.join(OUT, "gsdf_ref_colorize.so")
STAMP = os.path.join(OUT, ".stamp_colorize")


def main() -> int:
    if not os.path.exists(SRC):
        print(f"build_ref_colorize: {REF} not present; keeping prebuilt {SO} (if any)")
        return 0
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    for f in (SRC, WRAP, os.path.abspath(__file__)):
        h.update(open(f, "rb").read())
    stamp = h.hexdigest()
    if os.path.exists(SO) and os.path.exists(STAMP) and open(STAMP).read().strip() == stamp:
        print("build_ref_colorize: up to date")
        return 0
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}", f"-I{REF}"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    common = ["-std=c++17", "-O3", abi, "-DTORCH_EXTENSION_NAME=gsdf_ref_colorize", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    o1, o2 = os.path.join(OUT, "colorize_ref.o"), os.path.join(OUT, "colorize_wrap.o")
    p1 = subprocess.Popen(["nvcc", *common, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-w",
                           *inc, "-c", SRC, "-o", o1])
    p2 = subprocess.Popen(["g++", *common, "-fPIC", "-w", *inc, "-c", WRAP, "-o", o2])
    if p1.wait() or p2.wait():
        return 1
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    rc = subprocess.call(["g++", "-shared", "-o", SO, o1, o2, f"-L{libdir}", "-lc10", "-ltorch", "-ltorch_cpu",
                          "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-L/usr/local/cuda/lib64", "-lcudart",
                          f"-Wl,-rpath,{libdir}"])
    for o in (o1, o2):
        if os.path.exists(o):
            os.remove(o)
    if rc:
        return rc
    open(STAMP, "w").write(stamp)
    print(f"build_ref_colorize: built {SO}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
