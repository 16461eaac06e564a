"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT
from sqz_b200 import _lib


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sqz_[a-z0-9_]+)\s*\(", text)))


def test_headers_declare_what_the_binding_lists():
    names = set(declared("sqz.h")) | set(declared("sqz_gpu.h"))
    assert names == set(_lib.SYMBOLS), names ^ set(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    for name in declared("sqz.h") + declared("sqz_gpu.h"):
        assert getattr(L, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(_lib.SYMBOLS) <= exported
    assert all(s.startswith("sqz_") for s in exported), exported   # nothing else leaks


def test_headers_compile_as_c99_and_cxx(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "sqz.h"\n#include "sqz_gpu.h"\nint main(void){struct sqz_bitstream b; (void)b; return SQZ_GPU_ABI_VERSION - 1;}\n')
    inc = "-I" + os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", inc, "-c", str(src), "-o", str(tmp_path / "t.o")])
    cpp = tmp_path / "t.cpp"
    cpp.write_text(src.read_text())
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", inc, "-c", str(cpp), "-o", str(tmp_path / "t2.o")])


def test_struct_layout_matches_the_header(tmp_path):
    """ctypes mirrors of struct sqz / sqz_bitstream have the C sizes."""
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "sqz.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(struct sqz), '
                   'sizeof(struct sqz_bitstream), sizeof(struct sqz_tree));return 0;}\n')
    exe = tmp_path / "s"
    subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c = map(int, subprocess.check_output([str(exe)]).split())
    assert (a, b, c) == (C.sizeof(_lib.State), C.sizeof(_lib.Bitstream), C.sizeof(_lib.Tree))


def test_version_and_device_probe():
    L = _lib.load()
    assert L.sqz_gpu_abi_version() == 3
    assert L.sqz_gpu_device_count() >= 0
    assert L.sqz_gpu_parse_workspace(1 << 20) > 0
    assert L.sqz_gpu_select_kernel(7) != 0 and L.sqz_gpu_select_kernel(0) == 0


def test_product_never_links_the_oracle():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "sqzref" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "sqz_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_workspaces_cover_every_smaller_shard():
    """A workspace sized for n serves every shard of at most n positions (the pipeline sizes its slots
    once and sees chunks of many sizes; small shards use more slice tables and smaller parse blocks)."""
    import random
    L = _lib.load()
    rnd = random.Random(5)
    ns = sorted(set([1, 31, 32, 3968, 3969, 16256, 10**6, 1 << 20, 4 << 20, 5900000, 7217663, 7217665, 8 << 20, (8 << 20) + 1,
                     16 << 20, 32 << 20, 1 << 30] + [rnd.randrange(1, 40 << 20) for _ in range(200)]))
    pm = pp = 0
    for n in ns:
        m, p = L.sqz_gpu_match_workspace(n), L.sqz_gpu_parse_workspace(n)
        assert m >= pm and p >= pp, n
        pm, pp = m, p
