"""Decoder timings on the GPU box: host sqz_decompress vs sqz_decompress_gpu (host token reader + GPU copy phase)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqz_b200 as sq
from sqz_b200 import corpus
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d = corpus.synthetic(mb << 20, 0)
c = sq.compress(d, 15)
for it in range(2):
    t0 = time.perf_counter(); a = sq.decompress(c); t_host = time.perf_counter() - t0
    st = {}
    t0 = time.perf_counter(); b = sq.decompress_gpu(c, stats=st); t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter(); toks = sq.decode_tokens(c); t_tok = time.perf_counter() - t0
    t0 = time.perf_counter(); e = sq.expand_tokens(toks, d.size); t_exp = time.perf_counter() - t0
print("%d MiB, %d tokens: sqz_decompress %.2f s = %.1f MB/s; sqz_decompress_gpu %.2f s = %.1f MB/s "
      "(token reader %.2f s, expand incl. copies and allocation %.2f s); expand_tokens alone %.3f s; identical %s"
      % (mb, toks.size, t_host, d.size / 1e6 / t_host, t_gpu, d.size / 1e6 / t_gpu, st["entropy_seconds"],
         st["expand_seconds"], t_exp, a == b == e == d.tobytes()))
