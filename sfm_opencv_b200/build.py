"""In-tree nvcc build of libsfm_b200.so (sm_100a only; no other backend)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsfm_b200.so")
SOURCES = ["match_knn.cu", "match_hamming.cu", "match_hamming_tc.cu", "match_finalize.cu", "geometry.cu", "host_io.cu", "capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-cudart", "shared",
]
# The CUDA runtime is linked as a shared library (libcudart.so.12 of the image, or the copy a host
# process such as torch has already loaded): the library then carries none of the runtime's own
# symbol table, and shares streams / the primary context with its host process.
LINK_FLAGS = ["-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared",
              "-Xlinker", "-rpath=/usr/local/cuda/lib64"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsfm_b200.so cannot be built (there is no CPU fallback)")


def _stale(out: str, deps: list[str]) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "sfm_b200.h"))
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src[:-3] + ".o")
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            subprocess.run(cmd, check=True)
        objs.append(o)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + LINK_FLAGS + ["-o", LIB] + objs + ["-ldl", "-lpthread", "-lrt"]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
