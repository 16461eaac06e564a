"""Debugging aid: match-table time on 16 MiB made of one kind of data (text / ELF / image)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sqz_b200 import _lib, corpus
L = _lib.load()
fx = corpus.fixtures()
kinds = {"text": np.concatenate([fx["confucius.txt"], fx["laozi.txt"]]), "elf": np.concatenate([fx["x64.elf"], fx["arm64.elf"]]),
         "image": np.concatenate([fx["mandrill.bmp"], fx["mandrill.png"]])}
size = int(sys.argv[1]) if len(sys.argv) > 1 else 16 << 20
cyc = torch.zeros((1 << 20) + 64, dtype=torch.int64, device="cuda")
L.sqz_gpu_debug_tile_cycles(cyc.data_ptr())
for name, base in kinds.items():
    data = np.tile(base, size // base.size + 1)[:size].copy()
    buf = torch.zeros(size + 1024, dtype=torch.uint8, device="cuda"); buf[:size] = torch.from_numpy(data).cuda()
    table = torch.empty(size, dtype=torch.int32, device="cuda")
    best = 1e9
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.sqz_gpu_match_table_device(buf.data_ptr(), 0, size, 0, 3, 257, 32767, table.data_ptr(), torch.cuda.current_stream().cuda_stream)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    c = cyc.cpu().numpy(); dbg = c[1 << 20:]; cyc.zero_()
    print("%-6s rc %d %.1f ms %.0f MB/s | per position (3 runs): survivors %.2f better %.2f fresh-ties %.2f rejects %.2f handed %.3f hash-only %.3f | slow iterations %.1f%% of %d thread-iterations | phase2: searched %.3f%% inherited %.3f%% steps/search %.0f verifies/search %.1f" % (
        name, rc, best, size / 1e3 / best, dbg[8] / 3 / size, dbg[9] / 3 / size, dbg[10] / 3 / size, dbg[11] / 3 / size, dbg[12] / 3 / size, dbg[14] / 3 / size,
        100.0 * dbg[13] / max(1, 3 * (size // 128 + 1) * 8192), 3 * (size // 128 + 1) * 8192, 100.0 * dbg[0] / 3 / size, 100.0 * dbg[5] / 3 / size, dbg[1] / max(dbg[0], 1), dbg[2] / max(dbg[0], 1)), flush=True)
    why = {16: "neighbour open", 17: "neighbour without a match", 18: "neighbour at max_len, byte differs", 19: "byte differs",
           20: "no neighbour in the shard", 21: "other", 22: "capped inheritance's search"}
    print("       phase 2 searches by reason (% of positions): " + ", ".join("%s %.3f" % (w, 100.0 * dbg[k] / 3 / size) for k, w in why.items()) +
          " | verify rounds/search %.1f" % (dbg[4] / max(dbg[0], 1)), flush=True)
