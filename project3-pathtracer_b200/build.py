"""Build libpt_b200.so (CUDA kernels + C ABI + host loader / image writer / compat shim) and the headless
driver `pt_render`, in-tree, for sm_100a.  No GPU is needed to build (nvcc cross-compiles)."""
import glob
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libpt_b200.so")
DRIVER = os.path.join(PKG, "pt_render")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false + IEEE sqrt/div (nvcc defaults): the arithmetic contract of csrc/pt_device.cuh
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the B200 path tracer cannot be built (there is no CPU fallback)")
    return exe


def _sources():
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    cpp = sorted(p for p in glob.glob(os.path.join(CSRC, "*.cpp")) if os.path.basename(p) != "pt_main.cpp")
    return cu, cpp


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    cu, cpp = _sources()
    hdr = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(
        os.path.join(ROOT, "include", "*.h"))
    built = False
    if force or _stale(LIB, cu + cpp + hdr + [os.path.abspath(__file__)]):
        cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-I", os.path.join(ROOT, "include"), "-shared", "-o", LIB] + cu + cpp + ["-lz"]
        subprocess.check_call(cmd)
        built = True
    main = os.path.join(CSRC, "pt_main.cpp")
    if os.path.exists(main) and (force or built or _stale(DRIVER, [main, LIB])):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), main, "-o", DRIVER,
                               "-L", PKG, "-lpt_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
