"""The measurement inputs are what DESIGN.md says they are."""
import hashlib

import numpy as np

from sqz_b200 import corpus, shard


def test_fixture_pack_matches_the_reference_blobs(golden):
    fx = corpus.fixtures()
    assert list(fx) == corpus.ORDER
    sizes = {"laozi.txt": 20760, "confucius.txt": 67735, "x64.elf": 926536, "arm64.elf": 847400,
             "mandrill.bmp": 786570, "mandrill.png": 627896}
    blobs = {"arm64.elf": "5bffdf72ee3b", "confucius.txt": "59a6ebb3c3cd", "laozi.txt": "c450477f44a1",
             "mandrill.bmp": "c97881ee6d63", "mandrill.png": "90c78819b2e5", "x64.elf": "0d861f26259c"}
    for n, d in fx.items():
        assert d.size == sizes[n] == golden[n]["bytes"]
        sha = hashlib.sha1(b"blob %d\0" % d.size + d.tobytes()).hexdigest()[:12]
        assert sha == blobs[n] == golden[n]["git_blob"]


def test_synthetic_stream_is_a_pure_function_of_the_offset():
    B = corpus.base().size
    assert B == 3276897
    a = corpus.synthetic(3 * B + 1000, 0)
    assert (a[:B] == corpus.base()).all()                              # repetition 0 is verbatim
    for off, n in [(0, 10), (B - 5, 10), (B + 12345, 70000), (2 * B - 1, B + 2), (3 * B, 1000)]:
        assert (corpus.synthetic(n, off) == a[off:off + n]).all()
    diff = int((a[B:2 * B] != a[:B]).sum())
    assert 0.9 * (B // 64) * 255 / 256 < diff <= B // 64               # ~1/64 of the bytes mutated
    assert hashlib.sha1(a[B:B + 65536].tobytes()).hexdigest()[:12] == \
        hashlib.sha1(corpus.synthetic(65536, B).tobytes()).hexdigest()[:12]


def test_shard_plan_covers_the_input_with_the_right_halos():
    for total, world in [(1 << 20, 1), (1 << 20, 2), ((1 << 20) + 7, 8), (1000, 4), (5, 8)]:
        sh = shard.plan(total, world, 32767, 257)
        assert sh[0].first == 0 and sh[-1].first + sh[-1].n == total
        for a, b in zip(sh, sh[1:]):
            assert a.first + a.n == b.first
        for s in sh:
            assert s.back == min(s.first, 32767) and s.ahead == min(total - s.first - s.n, 257)
            assert s.lo >= 0 and s.hi <= total


def test_chain_entries():
    m0 = np.arange(257, dtype=np.uint16)[::-1].copy()      # entry e -> 256 - e
    m1 = (np.arange(257, dtype=np.uint16) // 2).astype(np.uint16)
    assert shard.chain_entries([m0, m1]) == [0, 256, 128]
    assert shard.chain_entries([m0, m1], first_entry=6) == [6, 250, 125]
