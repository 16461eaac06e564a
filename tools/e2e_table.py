"""Time the host-buffer C-ABI calls (pinned buffers) on the bench corpus."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sqz_b200 import _lib, corpus
L = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
def pinned(nbytes, dtype):
    p = L.sqz_gpu_host_alloc(nbytes)
    return p, np.frombuffer((C.c_uint8 * nbytes).from_address(p), dtype=dtype)
p_in, h_in = pinned(n, np.uint8); h_in[:] = corpus.synthetic(n, 0)
p_len, h_len = pinned(2 * n, np.uint16); p_dist, h_dist = pinned(2 * n, np.uint16)
p_tok, h_tok = pinned(4 * n, np.uint32)
for it in range(3):
    t0 = time.perf_counter()
    rc = L.sqz_gpu_match_table(C.cast(p_in, _lib.u8p), n, 1 << 15, 3, 257, 32767, C.cast(p_len, _lib.u16p), C.cast(p_dist, _lib.u16p))
    dt = time.perf_counter() - t0
    print("match_table rc %d %.3f s %.1f MB/s" % (rc, dt, n / 1e6 / dt), flush=True)
for it in range(2):
    cnt = C.c_size_t()
    t0 = time.perf_counter()
    rc = L.sqz_gpu_tokens(C.cast(p_in, _lib.u8p), n, 1 << 15, 3, 257, 32767, C.cast(p_tok, _lib.u32p), n, C.byref(cnt))
    dt = time.perf_counter() - t0
    print("tokens rc %d %.3f s %.1f MB/s (%d tokens)" % (rc, dt, n / 1e6 / dt, cnt.value), flush=True)
