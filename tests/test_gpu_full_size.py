"""BASELINE.json config 5 at full size: the 1 GiB synthetic corpus.  Opt-in (SQZ_FULL_SIZE=1):
it needs ~25 GB of host memory and a few minutes, so the default `-m gpu` run skips it.
Result of the round-1 run is recorded in DESIGN.md section 7."""
import os
import time

import numpy as np
import pytest

import sqz_b200 as sq
from sqz_b200 import corpus

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SQZ_FULL_SIZE") != "1", reason="set SQZ_FULL_SIZE=1")]


def test_one_gib_table_tokens_and_round_trip(oracle, reference):
    n = 1 << 30
    d = corpus.synthetic(n, 0)
    t0 = time.time()
    ln, ds = sq.match_table(d)
    t_gpu = time.time() - t0
    t0 = time.time()
    oln, ods = oracle.match_table(d, 1 << 15, fast=True)           # oracle B, == oracle A on all fixtures
    t_cpu = time.time() - t0
    assert (ln == oln).all() and (ds == ods).all()
    t = sq.tokens(d)
    ot, end = oracle.tokens_from_table(d, oln, ods)
    assert end == n and t.size == ot.size and (t == ot).all()
    # sampled positions against the restated reference loop itself (oracle A)
    rng = np.random.default_rng(11)
    for i in rng.integers(0, n, 3000).tolist():
        assert oracle.best(d, i, 1 << 15) == (int(ln[i]), int(ds[i])), i
    # bitstream: byte-identical to the reference's own encoder on the same tokens, and it round-trips
    part = np.ascontiguousarray(d[: 128 << 20])
    comp = sq.compress(part, 15)
    assert reference.encode_tokens(sq.tokens(part), part.size, 15) == comp
    assert sq.decompress(comp) == part.tobytes()
    assert reference.decompress(comp) == part.tobytes()
    print("1 GiB: GPU table %.1f s, oracle B %.1f s, %d tokens" % (t_gpu, t_cpu, t.size))
