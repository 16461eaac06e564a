"""GPU parity: the CUDA path through the C-ABI against the oracle, the golden
digests made with the unmodified reference, and the reference decoder."""
import numpy as np
import pytest

import sqz_b200 as sq
from conftest import fnv

pytestmark = pytest.mark.gpu

# the reference's six test/ files + csrc.cat, the stand-in for BASELINE config 2's sqlite3.c
FILES = ["laozi.txt", "confucius.txt", "x64.elf", "arm64.elf", "mandrill.bmp", "mandrill.png", "csrc.cat"]
KATS = ["zeros4096", "pat1234x1024", "hello", "abc40", "lorem3", "empty", "one", "two", "aaa", "aaaa"]


@pytest.fixture(params=[2, 1], ids=["bitsliced", "per_position"])
def kernel(request):
    """Both match-table kernels must be exact; 0/auto is restored afterwards."""
    sq.select_kernel(request.param)
    yield request.param
    sq.select_kernel(0)


@pytest.mark.parametrize("name", KATS + FILES)
@pytest.mark.parametrize("wb", [10, 15])
def test_match_table_golden(name, wb, inputs, golden, oracle, kernel):
    """Full table == digest of the oracle's table (oracle pinned to the reference)."""
    d = inputs[name]
    ln, ds = sq.match_table(d, 1 << wb)
    g = golden[name]["win"][str(wb)]
    if d.size <= 70000:   # small enough for the brute-force oracle: element-wise report
        oln, ods = oracle.match_table(d, 1 << wb)
        bad = np.nonzero((ln != oln) | (ds != ods))[0]
        assert bad.size == 0, (name, wb, bad[:5], ln[bad[:5]], oln[bad[:5]], ds[bad[:5]], ods[bad[:5]])
    assert fnv(oracle, ln) == g["fnv_len"]
    assert fnv(oracle, ds) == g["fnv_dist"]


@pytest.mark.parametrize("name", KATS + FILES)
@pytest.mark.parametrize("wb", [10, 15])
def test_tokens_and_bitstream_golden(name, wb, inputs, golden, oracle):
    d = inputs[name]
    g = golden[name]["win"][str(wb)]
    t = sq.tokens(d, 1 << wb)
    assert t.size == g["tokens"]
    assert int((t >> 16 != 0).sum()) == g["matches"]
    assert fnv(oracle, t) == g["fnv_tokens"]
    comp = sq.compress(d, wb)
    assert len(comp) == g["compressed_bytes"]
    assert fnv(oracle, np.frombuffer(comp, np.uint8)) == g["fnv_mem"]
    assert sq.decompress(comp) == d.tobytes()


@pytest.mark.parametrize("name", ["hello", "laozi.txt", "arm64.elf", "csrc.cat"])
def test_reference_decoder_accepts_our_stream(name, inputs, reference):
    d = inputs[name]
    comp = sq.compress(d, 15)
    assert reference.decompress(comp) == d.tobytes()
    comp_f = sq.compress(d, 15, file_mode=True)
    # file mode = the same 64-bit words in host byte order
    a = np.frombuffer(comp, np.uint8).reshape(-1, 8)[:, ::-1].reshape(-1)
    assert a.tobytes() == comp_f


def test_config_2_far_max_len_repeats(inputs, golden, oracle, reference):
    """BASELINE config 2 (csrc.cat for sqlite3.c): 17 % of the positions reach max_len at a mean
    distance of ~30,600 -- the reference's far early-out, squeeze.h:353.  Full table against the
    brute-force oracle, element by element; token stream and bytes against the unmodified reference."""
    d = inputs["csrc.cat"]
    ln, ds = sq.match_table(d, 1 << 15)
    oln, ods = oracle.match_table(d, 1 << 15)              # oracle A: the restated loop itself
    bad = np.nonzero((ln != oln) | (ds != ods))[0]
    assert bad.size == 0, (bad[:5], ln[bad[:5]], oln[bad[:5]], ds[bad[:5]], ods[bad[:5]])
    far = (ln == 257) & (ds > 30000)
    assert far.sum() > 0.15 * d.size                       # the case this input is here for
    comp = sq.compress(d, 15)
    assert comp == reference.compress(d, 15)
    t = sq.tokens(d, 1 << 15)
    assert (t == reference.tokens(comp)).all()


def test_compress_with_every_coder_team(reference):
    """sqz_compress on 20 MiB of the bench corpus (several chunks of the GPU stream whatever the chunking:
    the default ramp for one and two coder threads, quarter-of-the-input chunks with a short start for
    crews): one thread, model + emitter, crews of 2, 3 and 7 emitters and the automatic choice all give
    the bytes of the unmodified reference's encoder on the same tokens; `into` writes the caller's buffer."""
    from sqz_b200 import corpus
    d = corpus.synthetic(20 << 20, 11 << 20)
    want = reference.encode_tokens(sq.tokens(d), d.size, 15)
    buf = np.empty(sq.capacity(d.size), np.uint8)
    for threads in (1, 2, 3, 4, 8, 0):
        st = {}
        got = sq.compress(d, 15, threads=threads, stats=st, into=buf)
        assert got.base is buf or got is buf or np.shares_memory(got, buf)
        assert got.tobytes() == want, threads
        assert st["tokens"] > 0 and st["matches"] > 0
    assert sq.decompress(want) == d.tobytes()
    with pytest.raises(ValueError):
        sq.compress(d, 15, into=np.empty(16, np.uint8))
