"""sqz_compress / sqz_decompress on 16, 64 and 256 MiB of the corpus: one coder thread, two (model + bit
packing) and crews (model + coder_threads - 1 emitters, the caller appends their segments); 0 = automatic."""
import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqz_b200 as sq
from sqz_b200 import corpus
for mb in (16, 64, 256):
    d = corpus.synthetic(mb << 20, 0)
    for threads in (1, 2, 3, 4, 6, 8, 0):
        for it in range(2):
            st = {}
            t0 = time.perf_counter(); c = sq.compress(d, 15, stats=st, threads=threads); t = time.perf_counter() - t0
        print(mb, "MiB, %d coder thread(s) ->" % threads, len(c), "bytes; %.2f s = %.1f MB/s; search %.2f s entropy %.2f s; tokens %d"
              % (t, d.size/1e6/t, st['search_seconds'], st['entropy_seconds'], st['tokens']))
t0 = time.perf_counter(); out = sq.decompress(c); t = time.perf_counter() - t0
print("decompress %.2f s = %.1f MB/s" % (t, len(out)/1e6/t), out == d.tobytes())
