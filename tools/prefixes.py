"""Corpus prefixes (1 MiB, 16 MiB, 256 MiB of the synthetic stream), one GPU: the device
leg (match table + parse, inputs resident, CUDA events) and the host-to-host call sqz_gpu_match_table.

    python tools/prefixes.py [out.json] [sizes in MiB ...]
"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sqz_b200 import _lib, corpus
L = _lib.load()
out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prefixes_r02.json"
sizes = [int(x) << 20 for x in sys.argv[2:]] or [1 << 20, 16 << 20, 256 << 20]
stream = torch.cuda.current_stream()
rows = []
for n in sizes:
    host = corpus.synthetic(n, 0)
    p_in = L.sqz_gpu_host_alloc(n); h_in = np.frombuffer((C.c_uint8 * n).from_address(p_in), np.uint8); h_in[:] = host
    p_len = L.sqz_gpu_host_alloc(2 * n); p_dist = L.sqz_gpu_host_alloc(2 * n)
    d = torch.zeros(n + 64, dtype=torch.uint8, device="cuda"); d[:n].copy_(torch.from_numpy(host))
    table = torch.empty(n, dtype=torch.int32, device="cuda"); toks = torch.empty(n + 4, dtype=torch.int32, device="cuda")
    work = torch.empty(L.sqz_gpu_parse_workspace(n), dtype=torch.uint8, device="cuda")
    mwork = torch.empty(L.sqz_gpu_match_workspace(n), dtype=torch.uint8, device="cuda")
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    best_dev, best_host = 1e9, 1e9
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        rc = L.sqz_gpu_match_table_device_ws(d.data_ptr(), 0, n, 0, 3, 257, 32767, table.data_ptr(), mwork.data_ptr(), stream.cuda_stream)
        rc |= L.sqz_gpu_parse_device(d.data_ptr(), table.data_ptr(), n, 0, 3, 257, toks.data_ptr(), n, work.data_ptr(), res.data_ptr(), stream.cuda_stream)
        e1.record(stream); torch.cuda.synchronize()
        assert rc == 0, L.sqz_gpu_last_error()
        if it: best_dev = min(best_dev, e0.elapsed_time(e1) * 1e-3)
    for it in range(4):
        t0 = time.perf_counter()
        rc = L.sqz_gpu_match_table(C.cast(p_in, _lib.u8p), n, 1 << 15, 3, 257, 32767, C.cast(p_len, _lib.u16p), C.cast(p_dist, _lib.u16p))
        dt = time.perf_counter() - t0
        assert rc == 0, L.sqz_gpu_last_error()
        if it: best_host = min(best_host, dt)
    rows.append({"bytes": n, "device_ms": best_dev * 1e3, "device_MBps": n / 1e6 / best_dev,
                 "host_to_host_ms": best_host * 1e3, "host_to_host_MBps": n / 1e6 / best_host, "tokens": int(res[0].item())})
    print("%5d MiB: device %.2f ms (%.1f MB/s), host to host %.2f ms (%.1f MB/s)"
          % (n >> 20, best_dev * 1e3, n / 1e6 / best_dev, best_host * 1e3, n / 1e6 / best_host), flush=True)
    L.sqz_gpu_host_free(p_in); L.sqz_gpu_host_free(p_len); L.sqz_gpu_host_free(p_dist)
    del d, table, toks, work, mwork
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(rows, open(out_path, "w"), indent=1)
