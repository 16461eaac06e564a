"""CPU model of what the bit-sliced match kernel's scalar path sees (no GPU needed).

For a sample of positions of one fixture it measures the run length against every distance and
replays the kernel's order of evaluation (groups of 128 distances; inside a group sh = 31..0 outer,
t = 0..3 inner, d = 32 (m0 + t) - sh) with the kernel's gate (need exact up to min_len + 3, ties let
through on positions whose best is fresh) and its hand-over rule.  Prints, per position: candidates
that pass the gate, improvements, ties on fresh positions (each a read of the table word in global
memory in the kernel), rejects, and the share of positions handed to phase 2 -- for the kernel as it
is and for variants (best run kept in 4 bits, other reject limits).

usage: python tools/model_scalar_path.py [fixture] [first position] [positions]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sqz_b200 import corpus

name = sys.argv[1] if len(sys.argv) > 1 else "x64.elf"
first = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
count = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
MIN_LEN, GATED, WINDOW_BITS, MAX_DIST = 3, 3, 32, 32767
data = corpus.fixtures()[name]
first = min(first, data.size - count - 300)
pos = np.arange(first, first + count)
dmax = min(MAX_DIST, first)                       # every sampled position can look this far back
d = np.arange(1, dmax + 1)
# run[p, d-1] = common prefix of data[p..] and data[p-d..], capped at WINDOW_BITS (= "beyond the window")
alive = np.ones((count, dmax), dtype=bool)
run = np.zeros((count, dmax), dtype=np.uint8)
for k in range(WINDOW_BITS):
    a = data[pos + k][:, None]
    b = data[(pos[:, None] + k) - d[None, :]]
    alive &= a == b
    run += alive
    if not alive.any():
        break

# the kernel's order of distances
order = []
for m0 in range(1, (dmax + 31) // 32 + 1, 4):
    group = [32 * (m0 + t) - sh for sh in range(31, -1, -1) for t in range(4)]
    order.append([x for x in group if 1 <= x <= dmax])


def replay(limit, best_bits, gated=GATED):
    """All sampled positions at once, one distance of the kernel's order at a time."""
    cap = (1 << best_bits) - 1                     # runs above this cannot be stored: hand over
    best = np.zeros(count, np.int32); best_d = np.zeros(count, np.int32); rejects = np.zeros(count, np.int32)
    closed = np.zeros(count, bool)
    tot = dict(passed=0, better=0, fresh_tie=0, reject=0, handed=0, beyond=0)
    for group in order:
        fresh = np.zeros(count, bool)
        for dd in group:
            x = run[:, dd - 1].astype(np.int32)
            need = np.maximum(MIN_LEN, np.minimum(np.where(fresh, best, best + 1), MIN_LEN + gated))
            passed = (x >= need) & ~closed
            if not passed.any():
                continue
            tot["passed"] += int(passed.sum())
            beyond = passed & ((x >= WINDOW_BITS) | (x > cap))
            better = passed & ~beyond & (x > best)
            tie = passed & ~beyond & ~better & (x == best) & fresh
            nearer = tie & (dd < best_d)
            reject = passed & ~beyond & ~better & ~tie
            tot["beyond"] += int(beyond.sum()); tot["better"] += int(better.sum())
            tot["fresh_tie"] += int(tie.sum()); tot["reject"] += int(reject.sum())
            best = np.where(better, x, best)
            best_d = np.where(better | nearer, dd, best_d)
            fresh |= better
            rejects = np.where(better, 0, rejects + (reject | (tie & ~nearer)))
            handed = ~beyond & (rejects > limit) & ~closed
            tot["handed"] += int(handed.sum())
            closed |= beyond | handed
    return {k: v / count for k, v in tot.items()}


print("%s, positions %d..%d, %d distances each" % (name, first, first + count, dmax))
for label, limit, bits, gated in (("kernel as it is (limit 6, 5 bits, 3 levels)", 6, 5, 3),
                                  ("best in 4 bits (runs >= 16 to phase 2)", 6, 4, 3),
                                  ("limit 1", 1, 5, 3), ("limit 3", 3, 5, 3), ("limit 14", 14, 5, 3),
                                  ("2 exact levels", 6, 5, 2), ("4 exact levels", 6, 5, 4), ("6 exact levels", 6, 5, 6)):
    o = replay(limit, bits, gated)
    print("%-42s passed %.2f  better %.2f  fresh ties %.2f  rejects %.2f | to phase 2: beyond %.3f + handed %.3f"
          % (label, o["passed"], o["better"], o["fresh_tie"], o["reject"], o["beyond"], o["handed"]))
