"""Debugging aid: run the token path on one fixture a few times (for ncu launch lists)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqz_b200 as sq
from sqz_b200 import corpus
d = corpus.fixtures()[sys.argv[1] if len(sys.argv) > 1 else "confucius.txt"]
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t0 = time.perf_counter(); t = sq.tokens(d); print("%.2f ms" % ((time.perf_counter() - t0) * 1e3), t.size)
