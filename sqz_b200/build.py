"""Build libsqz_b200.so in-tree (sqz_b200/lib/) for sm_100a with nvcc + gcc.

No JIT, no torch extension machinery: the library is a plain C-ABI shared
object (include/sqz.h, include/sqz_gpu.h) that C, C++ and ctypes callers load.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libsqz_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall",
    "-Xptxas", "-v",
]
CC_FLAGS = ["-std=gnu11", "-O2", "-fPIC", "-Wall", "-Wextra"]
EXPORT = "-DSQZ_EXPORT=__attribute__((visibility(\"default\")))"


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".c", ".h", ".cuh"))) + [
        os.path.join(ROOT, "include", "sqz.h"), os.path.join(ROOT, "include", "sqz_gpu.h")]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False, variant: str | None = None,
          defines: tuple[str, ...] = ()) -> str:
    """The product library, or -- variant="name", defines=("-DSQZ_DEBUG_COUNTERS", ...) -- an
    experimental build of the same ABI under tools/variants/ (loaded through SQZ_B200_LIB)."""
    out_dir, lib = OUT_DIR, LIB
    if variant:
        out_dir = os.path.join(ROOT, "tools", "variants", variant + ".build")
        lib = os.path.join(ROOT, "tools", "variants", variant + ".so")
    elif not force and not stale():
        return LIB
    os.makedirs(out_dir, exist_ok=True)
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
    objs = []
    log = []
    for f in sorted(os.listdir(CSRC)):
        src = os.path.join(CSRC, f)
        obj = os.path.join(out_dir, f + ".o")
        if f.endswith(".cu"):
            cmd = [_nvcc(), *NVCC_FLAGS, *defines, *inc, "-c", src, "-o", obj]
        elif f.endswith(".c"):
            cmd = ["gcc", *CC_FLAGS, *defines, *inc, "-c", src, "-o", obj]
        else:
            continue
        p = subprocess.run(cmd, capture_output=True, text=True)
        log.append(p.stderr)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("compile failed: " + " ".join(cmd))
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs,
           "-Xlinker", "--version-script=" + os.path.join(CSRC, "exports.map")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(out_dir, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        sys.stderr.write("\n".join(log))
    return lib


if __name__ == "__main__":
    # python -m sqz_b200.build [--force] [--variant NAME -DFLAG ...]
    name = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose=True, variant=name,
                defines=tuple(a for a in sys.argv[1:] if a.startswith("-D"))))
