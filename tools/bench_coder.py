"""Host entropy stage alone, no GPU: tokens of a corpus prefix from the oracle (cached under .scratch/),
then sqz_encode_symbols with one, two and more coder threads and sqz_decompress, best of three.

    python tools/bench_coder.py [MiB]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sqz_b200 as sq
from sqz_b200 import corpus
from oracle import Oracle

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = mib << 20
d = corpus.synthetic(n, 0)
cache = os.path.join(os.path.dirname(__file__), "..", ".scratch", "tokens_%dm.npy" % mib)
if os.path.exists(cache):
    toks = np.load(cache)
else:
    o = Oracle.get()
    ln, ds = o.match_table(d, 1 << 15, fast=True)
    toks, end = o.tokens_from_table(d, ln, ds)
    assert end == n
    os.makedirs(os.path.dirname(cache), exist_ok=True)
    np.save(cache, toks)
words = sq.symbols_of_tokens(toks)
ref = None
for threads in (1, 2, 3, 4, 6, 8):
    best = 1e9
    for it in range(int(os.environ.get("REPS", "3"))):
        t0 = time.perf_counter(); comp = sq.encode_symbols(words, n, 15, threads=threads); best = min(best, time.perf_counter() - t0)
    ref = ref or comp
    assert comp == ref
    print("%d MiB, %d tokens, %d coder thread(s): %.3f s = %.2f ns/token = %.0f MB/s of input; %d bytes out"
          % (mib, toks.size, threads, best, best * 1e9 / toks.size, n / 1e6 / best, len(comp)))
best = 1e9
for it in range(int(os.environ.get("REPS", "3"))):
    t0 = time.perf_counter(); out = sq.decompress(comp); best = min(best, time.perf_counter() - t0)
assert out == d.tobytes()
print("sqz_decompress: %.3f s = %.0f MB/s of output, %.2f ns/token" % (best, n / 1e6 / best, best * 1e9 / toks.size))
