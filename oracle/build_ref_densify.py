"""TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/gsdf_ref_densify.so: the reference's OWN densification code —
GaussianModel::{prunePoints, densificationPostfix, densifyAndSplit, densifyAndClone, densifyAndPrune, resetOpacity,
replaceTensorToOptimizer, addDensificationStats} and the four accessors they call — compiled from where it lies.
/root/reference/src/gaussian_model.cpp as a whole needs Eigen / Sophus / OpenCV / PCL (absent here), so this script
EXTRACTS the function bodies by name (brace matching) at build time into a generated translation unit under
oracle/_ref/ (git-ignored, never committed) that starts with oracle/ref_densify_stub.h and ends with the pybind11
front oracle/ref_densify_wrap.inc.  No-op when /root/reference is absent or the output is up to date."""
import hashlib
import os
import re
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GSDF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SRC = os.path.join(REF, "src", "gaussian_model.cpp")
HDR = os.path.join(REF, "include", "gaussian_model.h")
STUB = os.path.join(HERE, "ref_densify_stub.h")
WRAP = os.path.join(HERE, "ref_densify_wrap.inc")
SO = os.path.join(OUT, "gsdf_ref_densify.so")
STAMP = os.path.join(OUT, ".stamp_densify")
FUNCS = ["getScalingActivation", "getXYZ", "getOpacityActivation", "percentDense", "resetOpacity",
         "replaceTensorToOptimizer", "prunePoints", "densificationPostfix", "densifyAndSplit", "densifyAndClone",
         "densifyAndPrune", "addDensificationStats"]


def extract_function(text: str, name: str) -> str:
    """The definition `<ret> GaussianModel::<name>(...) {...}` verbatim (from the start of its line to the closing brace)."""
    m = re.search(r"^[^\n;{}]*\bGaussianModel::" + name + r"\s*\(", text, flags=re.M)
    if not m:
        raise RuntimeError(f"GaussianModel::{name} not found in {SRC}")
    i = text.index("{", m.end())
    depth, j = 0, i
    while True:
        c = text[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return text[m.start(): j + 1]


def extract_macro(text: str, name: str) -> str:
    m = re.search(r"^#define " + name + r"\b.*?(?<!\\)\n", text, flags=re.M | re.S)
    if not m:
        raise RuntimeError(f"macro {name} not found in {HDR}")
    return m.group(0)


def main() -> int:
    if not os.path.exists(SRC):
        print(f"build_ref_densify: {REF} not present; keeping prebuilt {SO} (if any)")
        return 0
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    for f in (SRC, HDR, STUB, WRAP, os.path.abspath(__file__)):
        h.update(open(f, "rb").read())
    stamp = h.hexdigest()
    if os.path.exists(SO) and os.path.exists(STAMP) and open(STAMP).read().strip() == stamp:
        print("build_ref_densify: up to date")
        return 0
    src, hdr = open(SRC).read(), open(HDR).read()
    stub = open(STUB).read().replace("@TENSORS_TO_VEC_MACRO@", extract_macro(hdr, "GAUSSIAN_MODEL_TENSORS_TO_VEC"))
    gen_h = os.path.join(OUT, "densify_stub_gen.h")
    gen_c = os.path.join(OUT, "densify_gen.cpp")
    open(gen_h, "w").write(stub)
    parts = ['#include "densify_stub_gen.h"\n']
    for name in FUNCS:
        parts.append(f"// ---- extracted verbatim: GaussianModel::{name} ----\n" + extract_function(src, name) + "\n")
    parts.append(open(WRAP).read())
    open(gen_c, "w").write("\n".join(parts))
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{OUT}", f"-I{os.path.join(REF, 'include')}"] + [f"-I{p}" for p in ce.include_paths()] + \
          [f"-I{sysconfig.get_paths()['include']}"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    rc = subprocess.call(["g++", "-std=c++17", "-O2", abi, "-DTORCH_EXTENSION_NAME=gsdf_ref_densify",
                          "-DTORCH_API_INCLUDE_EXTENSION_H", "-fPIC", "-w", "-shared", *inc, gen_c, "-o", SO,
                          f"-L{libdir}", "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", f"-Wl,-rpath,{libdir}"])
    for f in (gen_h, gen_c):     # the generated files hold reference code: they never outlive the build
        if os.path.exists(f):
            os.remove(f)
    if rc:
        return rc
    open(STAMP, "w").write(stamp)
    print(f"build_ref_densify: built {SO}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
