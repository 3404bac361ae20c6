"""Builds q1tsim_b200/lib/libq1tsim.so (CUDA kernels + C++ host engine + C ABI)
for sm_100a with nvcc.  In-tree so that the .so travels to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libq1tsim.so")
SOURCES = ["kernels.cu", "engine.cu", "planner.cpp", "sampling.cpp", "gates.cpp", "capi.cpp", "circuit.cpp", "sharded.cpp", "composite.cpp", "export.cpp", "latex.cpp", "ffi_compat.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-ccbin", "g++",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-cudart", "static"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant="", defines=()):
    """variant: suffix for an A/B build (lib/libq1tsim<variant>.so) compiled with extra -D defines"""
    global LIB
    os.makedirs(LIBDIR, exist_ok=True)
    LIB = os.path.join(LIBDIR, "libq1tsim%s.so" % variant)
    objdir = os.path.join(LIBDIR, "obj" + variant)
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "q1t_engine.h"))
    headers.append(os.path.join(os.path.dirname(HERE), "include", "q1tsim_ffi.h"))
    headers = [h for h in headers if os.path.exists(h)]
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(objdir, src + ".o")
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            cmd = ["nvcc"] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-x", "cu", "-c", sp, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = ["nvcc", "-shared", "-ccbin", "g++", "-cudart", "static", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose=True, variant=var[0] if var else "", defines=defs))
