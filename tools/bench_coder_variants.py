"""A/B of host coder builds on the same tokens, no GPU: every C file given (a copy of sqz_codec.c, with
optional -D flags after a colon) is built with the GPU entry points stubbed and timed with 1, 2 and 4
coder threads and as a decoder, best of REPS.

    python tools/bench_coder_variants.py MiB file.c[:-DFLAG...] ...
"""
import os, sys, time, subprocess, tempfile, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sqz_b200 as sq
from sqz_b200 import corpus, _lib
from oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mib = int(sys.argv[1])
n = mib << 20
d = corpus.synthetic(n, 0)
o = Oracle.get()
ln, ds = o.match_table(d, 1 << 15, fast=True)
toks, end = o.tokens_from_table(d, ln, ds)
words = sq.symbols_of_tokens(toks)
STUB = ('#include "sqz_gpu.h"\n#include <errno.h>\n'
        'int sqz_gpu_stream_open(sqz_gpu_stream** s, int dev, const uint8_t* p, size_t n, uint32_t w, uint32_t a, uint32_t b, uint32_t c, size_t k, uint32_t m) { return ENODEV; }\n'
        'int sqz_gpu_stream_next(sqz_gpu_stream* s, const uint32_t** t, size_t* c) { return ENODEV; }\n'
        'void sqz_gpu_stream_close(sqz_gpu_stream* s) { }\n'
        'int sqz_gpu_expand_tokens(const uint32_t* t, size_t n, uint8_t* o, size_t b) { return ENODEV; }\n')
tmp = tempfile.mkdtemp()
open(os.path.join(tmp, "stub.c"), "w").write(STUB)
libs = []
for spec in sys.argv[2:]:
    src, _, flags = spec.partition(":")
    so = os.path.join(tmp, "v%d.so" % len(libs))
    subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-pthread", "-w", *flags.split(), "-I" + os.path.join(ROOT, "include"),
                           src, os.path.join(tmp, "stub.c"), "-o", so])
    L = C.CDLL(so)
    for fn_name in ("sqz_write_header", "sqz_init", "sqz_encode_tokens", "sqz_encode_symbols", "sqz_decompress_buffer"):
        fn = getattr(L, fn_name)
        fn.restype, fn.argtypes = _lib.SYMBOLS[fn_name]
    libs.append((spec, L))
reps = int(os.environ.get("REPS", "5"))
ref = None
best = {}
for it in range(reps):
    for spec, L in libs:
        for threads in (1, 2, 4):
            t0 = time.perf_counter(); comp = sq.encode_symbols(words, n, 15, threads=threads, lib=L); t = time.perf_counter() - t0
            ref = ref or comp
            assert comp == ref
            best[(spec, threads)] = min(best.get((spec, threads), 1e9), t)
for (spec, threads), t in best.items():
    print("%-60s %d thread(s): %.2f ns/token = %.0f MB/s of input" % (spec, threads, t * 1e9 / toks.size, n / 1e6 / t))
c = np.frombuffer(ref, np.uint8)
out = np.empty(n, np.uint8)
for spec, L in libs:
    t_best = 1e9
    for it in range(reps):
        got = C.c_uint64()
        t0 = time.perf_counter()
        rc = L.sqz_decompress_buffer(c.ctypes.data_as(_lib.u8p), c.size, out.ctypes.data_as(_lib.u8p), out.size, C.byref(got))
        t_best = min(t_best, time.perf_counter() - t0)
        assert rc == 0 and got.value == n
    assert out.tobytes() == d.tobytes()
    print("%-60s sqz_decompress: %.2f ns/token = %.0f MB/s of output" % (spec, t_best * 1e9 / toks.size, n / 1e6 / t_best))
