"""the reference CPU codec on a 16 MiB prefix of the synthetic corpus, both builds
(-O2 with asserts on, -O3 -DNDEBUG), one core each (run concurrently on two cores)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import Reference
from sqz_b200 import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 << 20
d = corpus.synthetic(n, 0)
out = {}
def run(release):
    r = Reference(release=release)
    c = r.compress(d, 15)
    out[release] = (r.last_seconds, len(c))
th = [threading.Thread(target=run, args=(rel,)) for rel in (False, True)]
[t.start() for t in th]; [t.join() for t in th]
for rel in (False, True):
    s, size = out[rel]
    print("%s: %d bytes -> %d in %.1f s = %.4f MB/s (1 core); extrapolated 1 GiB: %.1f h"
          % ("-O3 -DNDEBUG" if rel else "-O2, asserts on", n, size, s, n / 1e6 / s, s * (2**30 / n) / 3600), flush=True)
