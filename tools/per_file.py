"""BASELINE.json configs 1-4: every file of the reference's test/ directory and csrc.cat (the
stand-in for config 2's sqlite3.c), on 1, 2, 4 and 8 GPUs where the box has them.
Per file: GPU search+parse through the host C-ABI (sqz_gpu_tokens; sqz_gpu_tokens_multi for
more than one GPU), the whole codec (sqz_compress), and the unmodified reference's
squeeze.compress on one host core.

    python tools/per_file.py [out.json]        SQZ_SKIP_REF=1 skips the reference timing
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sqz_b200 as sq
from sqz_b200 import corpus
from oracle import Oracle, Reference

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/per_file_r02.json"
ref = Reference.get(release=True)
o = Oracle.get()
golden = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "golden.json")))["inputs"]
n_dev = sq.device_count()
worlds = [w for w in (1, 2, 4, 8) if w <= n_dev]
rows = []
files = corpus.all_files()
for w in worlds:                                    # warm up: contexts, kernels, parked buffers
    sq.tokens_multi(files["laozi.txt"], list(range(w)))
sq.tokens(files["laozi.txt"])


def best_of(fn, k=5):
    best = 1e9
    for _ in range(k):
        t0 = time.perf_counter(); r = fn(); best = min(best, time.perf_counter() - t0)
    return best, r


for name, d in files.items():
    t_tok, t = best_of(lambda: sq.tokens(d))
    t_c, comp = best_of(lambda: sq.compress(d, 15), 3)
    ok = "%016x" % o.fnv(np.frombuffer(comp, np.uint8)) == golden[name]["win"]["15"]["fnv_mem"]
    multi = {}
    for w in worlds:
        t_m, tm = best_of(lambda: sq.tokens_multi(d, list(range(w))))
        ok = ok and tm.size == t.size and bool((tm == t).all())
        multi[str(w)] = t_m
    if os.environ.get("SQZ_SKIP_REF") == "1":          # quick runs: keep the recorded reference times
        t_ref, same = float("nan"), True
    else:
        rc = ref.compress(d, 15); t_ref = ref.last_seconds
        same = rc == comp
    rows.append({"file": name, "bytes": int(d.size), "compressed": len(comp), "tokens": int(t.size),
                 "gpu_tokens_s": t_tok, "gpu_tokens_MBps": d.size / 1e6 / t_tok,
                 "gpu_tokens_multi_s": multi, "gpu_tokens_multi_MBps": {k: d.size / 1e6 / v for k, v in multi.items()},
                 "sqz_compress_s": t_c, "sqz_compress_MBps": d.size / 1e6 / t_c,
                 "reference_s": t_ref, "reference_MBps": d.size / 1e6 / t_ref, "identical": bool(ok and same)})
    print("%-14s %8d B -> %7d B, %7d tokens | GPU search+parse %7.2f ms (%6.1f MB/s) | %s | sqz_compress %7.1f ms (%5.1f MB/s) | "
          "reference %6.2f s (%.4f MB/s) | identical: %s"
          % (name, d.size, len(comp), t.size, t_tok * 1e3, d.size / 1e6 / t_tok,
             " ".join("%dgpu %.2f ms" % (int(k), v * 1e3) for k, v in multi.items()),
             t_c * 1e3, d.size / 1e6 / t_c, t_ref, d.size / 1e6 / t_ref, ok and same), flush=True)
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(rows, open(out_path, "w"), indent=1)
