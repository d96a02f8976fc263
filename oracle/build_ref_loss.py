"""TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/gsdf_ref_loss.so: the reference's UNMODIFIED include/loss_utils.h
(a LibTorch-only header, included where it lies under /root/reference, never copied) behind the pybind11 wrapper
oracle/ref_loss.cpp, against this image's LibTorch.  No-op when /root/reference is absent (the GPU box uses the
prebuilt file) or when the output is up to date."""
import hashlib
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GSDF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
HDR = os.path.join(REF, "include", "loss_utils.h")
WRAP = os.path.join(HERE, "ref_loss.cpp")
SO = os.path.join(OUT, "gsdf_ref_loss.so")
STAMP = os.path.join(OUT, ".stamp_loss")


def main() -> int:
    if not os.path.exists(HDR):
        print(f"build_ref_loss: {REF} not present; keeping prebuilt {SO} (if any)")
        return 0
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    for f in (HDR, WRAP, os.path.abspath(__file__)):
        h.update(open(f, "rb").read())
    stamp = h.hexdigest()
    if os.path.exists(SO) and os.path.exists(STAMP) and open(STAMP).read().strip() == stamp:
        print("build_ref_loss: up to date")
        return 0
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}",
                                                    f"-I{os.path.join(REF, 'include')}"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    rc = subprocess.call(["g++", "-std=c++17", "-O2", abi, "-DTORCH_EXTENSION_NAME=gsdf_ref_loss",
                          "-DTORCH_API_INCLUDE_EXTENSION_H", "-fPIC", "-w", "-shared", *inc, WRAP, "-o", SO,
                          f"-L{libdir}", "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", f"-Wl,-rpath,{libdir}"])
    if rc:
        return rc
    open(STAMP, "w").write(stamp)
    print(f"build_ref_loss: built {SO}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
