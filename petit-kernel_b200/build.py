"""In-tree build of libpetit_b200.so (nvcc, sm_100a, no torch) and of the torch
extension petit_kernel/ops*.so (g++, links libpetit_b200.so via $ORIGIN rpath).

Replaces the reference's CMake/Hunter build (CMakeLists.txt, setup.py:25-35), which
needs network access.  Usage: python build.py [--force] [--no-torch]
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
PKG = os.path.join(HERE, "petit_kernel")
LIB = os.path.join(PKG, "libpetit_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = ["fp4_gemm.cu", "repack.cu", "capi.cu", "allreduce.cu"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + cmd[0])
    if r.stderr.strip() and os.environ.get("PETIT_BUILD_VERBOSE"):
        sys.stderr.write(r.stderr)


def headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(ROOT, "include", "causalflow", "petit", "petit.h"))
    return hs


def build_lib(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hs = headers()
    # objects built with other flags (e.g. -DPETIT_DEBUG_HOOKS experiments) must not survive
    flags = os.environ.get("PETIT_EXTRA_NVCC_FLAGS", "")
    stamp = os.path.join(OBJ, ".flags")
    old = open(stamp).read() if os.path.exists(stamp) else None
    if old != flags:
        force = True
        with open(stamp, "w") as f:
            f.write(flags)

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _newer(obj, [os.path.join(CSRC, src)] + hs):
            # PETIT_EXTRA_NVCC_FLAGS: e.g. -DPETIT_DEBUG_HOOKS for the experiment switches
            _run([NVCC, *ARCH, *os.environ.get("PETIT_EXTRA_NVCC_FLAGS", "").split(),
                  "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                  "-Xptxas", "-v" if os.environ.get("PETIT_BUILD_VERBOSE") else "-O3",
                  "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                  "-c", os.path.join(CSRC, src), "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=len(CU_SOURCES)) as ex:
        objs = list(ex.map(compile_one, CU_SOURCES))
    if force or _newer(LIB, objs):
        _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"])
    return LIB


def build_torch_ext(force: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce

    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(PKG, "ops" + suffix)
    src = os.path.join(CSRC, "pybind.cc")
    if not (force or _newer(out, [src, LIB] + headers())):
        return out
    inc = []
    for p in ce.include_paths() + [sysconfig.get_paths()["include"],
                                   os.path.join(ROOT, "include"), CSRC,
                                   "/usr/local/cuda/include"]:
        inc += ["-I", p]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w",
           f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_EXTENSION_NAME=ops",
           "-DTORCH_API_INCLUDE_EXTENSION_H", *inc, src, "-o", out,
           "-L", PKG, "-lpetit_b200", "-L", torch_lib,
           "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
           "-L", "/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"]
    _run(cmd)
    return out


def main() -> None:
    force = "--force" in sys.argv
    print("built", build_lib(force))
    if "--no-torch" not in sys.argv:
        print("built", build_torch_ext(force))


if __name__ == "__main__":
    main()
