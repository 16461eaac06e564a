"""Device time of the decoder's copy phase (lz_expand.cuh) with tokens and output resident in HBM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqz_b200 as sq
from sqz_b200 import _lib, corpus
L = _lib.load()
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = corpus.synthetic(mb << 20, 0)
toks = sq.tokens(d)
d_tok = torch.from_numpy(toks.view(np.int32)).cuda()
out = torch.empty(d.size, dtype=torch.uint8, device="cuda")
work = torch.empty(L.sqz_gpu_expand_workspace(toks.size, d.size), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
before = sq.launch_count()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = L.sqz_gpu_expand_tokens_device(d_tok.data_ptr(), toks.size, out.data_ptr(), d.size, work.data_ptr(), s)
    e1.record(); torch.cuda.synchronize()
    assert rc == 0, L.sqz_gpu_last_error()
    ms = e0.elapsed_time(e1)
launches = (sq.launch_count() - before) // 3
rounds = launches - 4
same = bool((out.cpu().numpy() == d).all())
alg = toks.size * 4 * 2 + d.size * (4 + 1) + rounds * d.size * 12 + d.size * 5      # scan + place, doubling rounds, fetch
print("%d MiB, %d tokens: %.2f ms = %.1f GB/s of output, %d doubling rounds, at most %.0f GB/s of HBM traffic by the algorithm's count (every round counted in full); identical %s"
      % (mb, toks.size, ms, d.size / 1e6 / ms, rounds, alg / 1e6 / ms, same))
