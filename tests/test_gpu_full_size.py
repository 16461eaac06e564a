"""BASELINE.json config 5 at full size: the 1 GiB synthetic corpus, part of the default `-m gpu`
run.  It needs ~25 GB of host memory and two to three minutes (most of it oracle B on the host
cores); on a host with less than 40 GB available it runs on the first 256 MiB instead, and
SQZ_FULL_SIZE=0 forces that size."""
import os
import time

import numpy as np
import pytest

import sqz_b200 as sq
from sqz_b200 import corpus

pytestmark = pytest.mark.gpu


def _size() -> int:
    if os.environ.get("SQZ_FULL_SIZE") == "0":
        return 256 << 20
    try:
        import psutil
        if psutil.virtual_memory().available < (40 << 30):
            return 256 << 20
    except ImportError:
        pass
    return 1 << 30


def test_full_size_table_tokens_and_round_trip(oracle, reference):
    n = _size()
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
    del ot
    # the multi-device entry point, the input cut in four (on one device when the box has no more)
    import torch
    devices = [g % torch.cuda.device_count() for g in range(4)]
    tm = sq.tokens_multi(d, devices)
    assert tm.size == t.size and (tm == t).all()
    del tm
    # sampled positions against the restated reference loop itself (oracle A)
    rng = np.random.default_rng(11)
    for i in rng.integers(0, n, 1000).tolist():
        assert oracle.best(d, i, 1 << 15) == (int(ln[i]), int(ds[i])), i
    # bitstream: byte-identical to the reference's own encoder on the same tokens, and it round-trips
    part = np.ascontiguousarray(d[: 64 << 20])
    comp = sq.compress(part, 15)
    assert reference.encode_tokens(sq.tokens(part), part.size, 15) == comp
    assert sq.decompress(comp) == part.tobytes()
    assert reference.decompress(comp) == part.tobytes()
    print("%d MiB: GPU table %.1f s, oracle B %.1f s, %d tokens" % (n >> 20, t_gpu, t_cpu, t.size))
