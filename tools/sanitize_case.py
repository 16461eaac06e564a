"""Smallest set of calls that touches every kernel and every edge path; meant for
`compute-sanitizer --tool memcheck python tools/sanitize_case.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqz_b200 as sq
from sqz_b200 import corpus, device
fx = corpus.fixtures()
d = np.concatenate([fx["laozi.txt"], fx["x64.elf"][905000:925000]])      # text + zero-rich ELF tail, 3 tiles
for wb in (10, 15):
    ln, ds = sq.match_table(d, 1 << wb)
    t = sq.tokens(d, 1 << wb)
    print("window 2^%d: %d positions, %d tokens, %d matches" % (wb, d.size, t.size, int((ln > 0).sum())))
ln, ds = sq.match_table(d, 1 << 15, 2, 254, 32768)                       # rule set iii, min_len 2
ln, ds = sq.match_table(d[:5000], 1 << 12, 4, 16, None)                  # thread-per-position kernel
# a shard with halos at an odd address, interior + partial tiles
total = 120000
data = corpus.synthetic(total, 3276897 - 60000)
buf = torch.zeros(total + 3 + 64, dtype=torch.uint8, device="cuda")
buf[3:3 + total] = torch.from_numpy(data).cuda()
tab = device.match_table(buf, 3 + 40000, 50001, 32767, 257)
em = device.exit_map(tab, 50001)
tok, over = device.parse(buf, 3 + 40000, tab, 50001, 0)
torch.cuda.synchronize()
print("shard ok", int(tok.numel()), over)
comp = sq.compress(d, 15)
assert sq.decompress(comp) == d.tobytes()
print("done")
