"""Tuning aid (library built with -DSQZ_TUNING): time sqz_gpu_tokens on small inputs for forced tile
shapes (SQZ_Q = blocks per thread) and slice counts (SQZ_SLICES), against the automatic choice."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sqz_b200 as sq
from sqz_b200 import corpus
files = corpus.all_files()
inputs = {"laozi.txt": files["laozi.txt"], "confucius.txt": files["confucius.txt"], "csrc.cat": files["csrc.cat"],
          "x64.elf": files["x64.elf"], "mandrill.bmp": files["mandrill.bmp"],
          "1MiB": corpus.synthetic(1 << 20, 0), "4MiB": corpus.synthetic(4 << 20, 0), "6MiB": corpus.synthetic(6 << 20, 0)}
ref = {k: sq.tokens(v) for k, v in inputs.items()}
configs = [(0, 0)] + [(q, s) for q in (1, 4) for s in (1, 2, 4, 8, 16, 32)]
print("%-14s" % "q,slices" + " ".join("%9s" % k for k in inputs))
for q, s in configs:
    os.environ.pop("SQZ_Q", None); os.environ.pop("SQZ_SLICES", None)
    if q: os.environ["SQZ_Q"] = str(q); os.environ["SQZ_SLICES"] = str(s)
    row = []
    for k, d in inputs.items():
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter(); t = sq.tokens(d); best = min(best, time.perf_counter() - t0)
        assert t.size == ref[k].size and (t == ref[k]).all(), (q, s, k)
        row.append(best * 1e3)
    print("%-14s" % ("auto" if not q else "q%d s%d" % (q, s)) + " ".join("%9.2f" % x for x in row), flush=True)
