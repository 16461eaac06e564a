"""How fast the GPU stream alone delivers symbol words (no entropy coding): sqz_gpu_stream_open/next/close
over a corpus prefix with the default chunking (2, 4, 8, 16, then 32 MiB) and with fixed chunk sizes.

    python tools/stream_rate.py [MiB]
"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sqz_b200 import corpus, _lib

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = corpus.synthetic(mib << 20, 0)
L = _lib.load()
for chunk_mib in (0, 8, 16, 32, 64, 128):
    best, per = 1e9, None
    for it in range(3):
        st = C.c_void_p()
        t0 = time.perf_counter()
        rc = L.sqz_gpu_stream_open(C.byref(st), -1, d.ctypes.data_as(_lib.u8p), d.size, 1 << 15, 3, 257, 32767,
                                   chunk_mib << 20, 1)
        assert rc == 0, rc
        stamps, total = [], 0
        while True:
            words, count = _lib.u32p(), C.c_size_t()
            rc = L.sqz_gpu_stream_next(st, C.byref(words), C.byref(count))
            assert rc == 0, rc
            if count.value == 0:
                break
            total += count.value
            stamps.append(time.perf_counter() - t0)
        L.sqz_gpu_stream_close(st)
        t = time.perf_counter() - t0
        if t < best:
            best, per = t, stamps
    gaps = [per[0]] + [b - a for a, b in zip(per, per[1:])]
    print("chunk %3s MiB: %d MiB in %.3f s = %.0f MB/s, %d tokens; chunks arrive after (ms): %s"
          % (chunk_mib or "dflt", mib, best, d.size / 1e6 / best, total, " ".join("%.0f" % (g * 1e3) for g in gaps[:12])))
