"""Debugging aid: per-tile duration of the bit-sliced match kernel on one input."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sqz_b200 import _lib, corpus
L = _lib.load()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 32 << 20
data = corpus.synthetic(size, 0)
d = torch.from_numpy(data).cuda()
pad = torch.zeros(size + 1024, dtype=torch.uint8, device="cuda"); pad[:size] = d
table = torch.empty(size, dtype=torch.int32, device="cuda")
TP = 4 * (32 * 4 - 1) * 32      # v2::kTilePos
tiles = (size + TP - 1) // TP
cyc = torch.zeros((1 << 20) + 64, dtype=torch.int64, device="cuda")
L.sqz_gpu_debug_tile_cycles(cyc.data_ptr())
for it in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = L.sqz_gpu_match_table_device(pad.data_ptr(), 0, size, 0, 3, 257, 32767, table.data_ptr(), torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
    print("rc", rc, "ms", e0.elapsed_time(e1))
call = cyc.cpu().numpy(); c = call[:tiles]; dbg = call[1 << 20:] // 2
print('search reasons: neighbour open %d, neighbour has no match %d, neighbour at max_len %d, byte differs %d, shard end %d, other %d' % tuple(int(x) for x in dbg[16:22]))
print('finish: searched %d (%.3f%%), inherited %d (%.3f%%), word-steps/search %.1f, verifies/search %.1f, improvements/search %.2f' % (dbg[0], 100.0*dbg[0]/size, dbg[5], 100.0*dbg[5]/size, dbg[1]/max(dbg[0],1), dbg[2]/max(dbg[0],1), dbg[3]/max(dbg[0],1)))
raw = call[1 << 20:]
surv, better, tie_fresh, reject, hand, it_slow = (int(x) // 2 for x in raw[8:14])
warp_iters = tiles * 4 * ((32767 + 127) // 128) * 32
print('phase 1 scalar path: survivors %d (%.2f per position), improvements %d, fresh ties %d, rejects %d, handed over %d; '
      'thread-iterations that entered it %d = %.2f%% of %d thread-iterations (%.1f%% of warp-iterations if spread evenly)'
      % (surv, surv / size, better, tie_fresh, reject, hand, it_slow, 100.0 * it_slow / (warp_iters * 32), warp_iters * 32,
         100.0 * min(1.0, it_slow / warp_iters)))
print("tiles", tiles, "sum Gcyc", c.sum() / 1e9, "median", np.median(c), "p90", np.percentile(c, 90), "max", c.max())
order = np.argsort(-c)[:12]
B = corpus.base().size
names = corpus.ORDER; sizes = [corpus.fixtures()[n].size for n in names]
def where(pos):
    o = pos % B
    for n, s in zip(names, sizes):
        if o < s: return "%s+%d" % (n, o)
        o -= s
for t in order:
    print(t, c[t], "Mcyc %.1f" % (c[t] / 1e6), where(t * TP))
np.save("gpurun_out/tile_cycles.npy", c)
