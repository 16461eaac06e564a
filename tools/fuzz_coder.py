"""Fuzz campaign (not part of the suite, no GPU): random token streams through every coder team (1, 2, 3
and 5 threads) of the product build, of a build with 512-token blocks and of the self-check build with
256-token blocks, against the unmodified reference's encoder (oracle/_ref) and decoder.

    [FUZZ_BASE=seed] python tools/fuzz_coder.py [streams]
"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, ctypes as C, subprocess, tempfile
import sqz_b200 as sq
from sqz_b200 import _lib
from oracle import Reference
from test_codec import _random_stream, _skewed
tmp = tempfile.mkdtemp()
stub = os.path.join(tmp, 'stub.c')
open(stub, 'w').write('#include "sqz_gpu.h"\n#include <errno.h>\nint sqz_gpu_stream_open(sqz_gpu_stream** s, int dev, const uint8_t* p, size_t n, uint32_t w, uint32_t a, uint32_t b, uint32_t c, size_t k, uint32_t m) { return ENODEV; }\nint sqz_gpu_stream_next(sqz_gpu_stream* s, const uint32_t** t, size_t* c) { return ENODEV; }\nvoid sqz_gpu_stream_close(sqz_gpu_stream* s) { }\nint sqz_gpu_expand_tokens(const uint32_t* t, size_t n, uint8_t* o, size_t b) { return ENODEV; }\n')
def variant(*flags):
    so = os.path.join(tmp, 'v%d.so' % abs(hash(flags)))
    subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-pthread", "-w", *flags, "-I" + ROOT + "/include", ROOT + "/sqz_b200/csrc/sqz_codec.c", stub, "-o", so])
    L = C.CDLL(so)
    for fn_name in ("sqz_write_header", "sqz_init", "sqz_encode_tokens", "sqz_encode_symbols", "sqz_decompress_buffer"):
        fn = getattr(L, fn_name); fn.restype, fn.argtypes = _lib.SYMBOLS[fn_name]
    return L
small = variant("-DSQZ_SELFCHECK", "-DSQZ_PART_TOKENS=16")
mid = variant("-DSQZ_PART_TOKENS=32")
ref = Reference.get()
n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 300
t0 = time.time(); tokens_total = 0
for seed in range(n_streams):
    rng = np.random.default_rng(int(os.environ.get("FUZZ_BASE", "90000")) + seed)
    if seed % 7 == 3:
        t = _skewed(int(rng.integers(1000, 80000)), seed); nbytes = int(np.where(t >> 16 != 0, t >> 16, 1).sum())
    else:
        t, nbytes = _random_stream(rng)
        if seed % 5 == 0:      # longer: repeat the stream's literals a few times so that blocks get going
            lit = t[t < 256]
            if lit.size > 100:
                t = np.concatenate([t] + [lit[rng.permutation(lit.size)] for _ in range(int(rng.integers(2, 12)))]).astype(np.uint32)
                nbytes = int(np.where(t >> 16 != 0, t >> 16, 1).sum())
    want = ref.encode_tokens(t, nbytes, 15)
    words = sq.symbols_of_tokens(t)
    for threads in (1, 2, 3, 5):
        assert sq.encode_symbols(words, nbytes, 15, threads=threads) == want, (seed, threads, "product")
        assert sq.encode_symbols(words, nbytes, 15, threads=threads, lib=mid) == want, (seed, threads, "mid")
    assert sq.encode_symbols(words, nbytes, 15, threads=1 + seed % 3, lib=small) == want, (seed, "small")
    assert sq.decompress(want) == ref.decompress(want), (seed, "decoder")
    tokens_total += t.size
    if seed % 50 == 49: print(seed + 1, "streams,", tokens_total, "tokens, %.0f s" % (time.time() - t0), flush=True)
print("fuzz ok:", n_streams, "streams,", tokens_total, "tokens")
