"""BASELINE.json configs 1-4: every fixture of the reference's test/ directory, one GPU.
Per file: GPU search+parse through the host C-ABI (sqz_gpu_tokens), the whole codec
(sqz_compress), and the unmodified reference's squeeze.compress on one host core."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sqz_b200 as sq
from sqz_b200 import corpus
from oracle import Oracle, Reference

ref = Reference.get(release=True)
o = Oracle.get()
golden = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "golden.json")))["inputs"]
rows = []
sq.tokens(corpus.fixtures()["laozi.txt"])          # warm up (context, parked buffers)
for name, d in corpus.fixtures().items():
    best_tok = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); t = sq.tokens(d); best_tok = min(best_tok, time.perf_counter() - t0)
    best_c = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); comp = sq.compress(d, 15); best_c = min(best_c, time.perf_counter() - t0)
    ok = "%016x" % o.fnv(np.frombuffer(comp, np.uint8)) == golden[name]["win"]["15"]["fnv_mem"]
    if os.environ.get("SQZ_SKIP_REF") == "1":          # quick runs: keep the recorded reference times
        t_ref, same = float("nan"), True
    else:
        rc = ref.compress(d, 15); t_ref = ref.last_seconds
        same = rc == comp
    rows.append((name, d.size, len(comp), t.size, best_tok, best_c, t_ref, ok and same))
    print("%-14s %8d B -> %7d B, %7d tokens | GPU search+parse %7.2f ms (%6.1f MB/s) | sqz_compress %7.1f ms (%5.1f MB/s) | "
          "reference %6.2f s (%.4f MB/s) | speed-up %6.0fx | identical: %s"
          % (name, d.size, len(comp), t.size, best_tok * 1e3, d.size / 1e6 / best_tok, best_c * 1e3, d.size / 1e6 / best_c,
             t_ref, d.size / 1e6 / t_ref, t_ref / best_c, ok and same), flush=True)
json.dump([dict(zip(["file", "bytes", "compressed", "tokens", "gpu_tokens_s", "sqz_compress_s", "reference_s", "identical"], r))
           for r in rows], open("gpurun_out/per_file_r01.json", "w"), indent=1)
