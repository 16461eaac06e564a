"""Host entropy stage (csrc/sqz_codec.c) without a GPU: bit-exact against the golden digests
made with the unmodified reference, round trips, header bytes, error behaviour."""
import ctypes as C
import errno
import os

import numpy as np
import pytest

import sqz_b200 as sq
from conftest import ROOT, fnv
from sqz_b200 import _lib

ALL = ["zeros4096", "pat1234x1024", "hello", "abc40", "lorem3", "empty", "one", "two", "aaa", "aaaa",
       "laozi.txt", "confucius.txt", "x64.elf", "arm64.elf", "mandrill.bmp", "mandrill.png"]


def oracle_tokens(oracle, d, wb):
    ln, ds = oracle.match_table(d, 1 << wb, fast=True)
    t, end = oracle.tokens_from_table(d, ln, ds)
    assert end == d.size
    return t


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("wb", [10, 15])
def test_bitstream_identical_to_reference(name, wb, inputs, golden, oracle):
    d = inputs[name]
    g = golden[name]["win"][str(wb)]
    t = oracle_tokens(oracle, d, wb)
    comp = sq.encode_tokens(t, d.size, wb)
    assert len(comp) == g["compressed_bytes"] and len(comp) % 8 == 0
    assert fnv(oracle, np.frombuffer(comp, np.uint8)) == g["fnv_mem"]
    if "hex_mem" in g:
        assert comp.hex() == g["hex_mem"]
    comp_f = sq.encode_tokens(t, d.size, wb, file_mode=True)      # callback mode: host-order words
    assert fnv(oracle, np.frombuffer(comp_f, np.uint8)) == g["fnv_file"]
    assert np.frombuffer(comp, np.uint8).reshape(-1, 8)[:, ::-1].tobytes() == comp_f
    assert sq.decompress(comp) == d.tobytes()
    assert sq.read_header(comp) == (d.size, wb)


@pytest.mark.parametrize("name", ["hello", "zeros4096", "laozi.txt", "confucius.txt"])
def test_cross_decoding_with_the_reference(name, inputs, reference, oracle):
    d = inputs[name]
    theirs = reference.compress(d, 15)
    assert sq.decompress(theirs) == d.tobytes()
    ours = sq.encode_tokens(oracle_tokens(oracle, d, 15), d.size, 15)
    assert ours == theirs
    assert reference.decompress(ours) == d.tobytes()


@pytest.mark.parametrize("name", ["hello", "abc40", "zeros4096", "laozi.txt", "arm64.elf", "mandrill.bmp"])
def test_symbol_words_code_to_the_same_bytes(name, inputs, golden, oracle):
    """sqz_encode_symbols (the form the GPU parse hands over) == sqz_encode_tokens."""
    d = inputs[name]
    t = oracle_tokens(oracle, d, 15)
    w = sq.symbols_of_tokens(t)
    comp = sq.encode_symbols(w, d.size, 15)
    assert fnv(oracle, np.frombuffer(comp, np.uint8)) == golden[name]["win"]["15"]["fnv_mem"]
    assert sq.encode_symbols(w, d.size, 15, file_mode=True) == sq.encode_tokens(t, d.size, 15, file_mode=True)


def test_symbol_word_layout():
    """bits 0..8 lit/len symbol, 9..13 length extra, 14..18 distance bucket, 19..31 distance extra;
    extra bits are stored in emission order (bit-reversed within the field): squeeze.h:29-79,290-315."""
    len_base = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227]
    len_xb = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4
    pos_base = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049,
                3073, 4097, 6145, 8193, 12289, 16385, 24577]
    pos_xb = [0, 0, 0, 0] + [k // 2 for k in range(2, 28)]
    rev = lambda v, n: int(format(v, "0%db" % n)[::-1], 2) if n else 0
    lens = np.arange(3, 258, dtype=np.uint32)
    dists = np.r_[np.arange(1, 600), np.arange(4090, 4100), np.arange(24570, 24580), 32766, 32767].astype(np.uint32)
    toks = np.r_[np.arange(256, dtype=np.uint32), (lens[:, None] << 16 | dists[None, :]).ravel()]
    words = sq.symbols_of_tokens(toks)
    assert (words[:256] == np.arange(256)).all()
    for t, w in zip(toks[256::37].tolist(), words[256::37].tolist()):
        ln, ds = t >> 16, t & 0xFFFF
        lb = max(k for k in range(28) if len_base[k] <= ln)
        pb = max(k for k in range(30) if pos_base[k] <= ds)
        assert w & 0x1FF == 257 + lb
        assert (w >> 9) & 31 == rev(ln - len_base[lb], len_xb[lb])
        assert (w >> 14) & 31 == pb
        assert w >> 19 == rev(ds - pos_base[pb], pos_xb[pb])
    bad = np.array([(258 << 16) | 1, (3 << 16) | 0x8000, (2 << 16) | 5, (3 << 16) | 0, 256], np.uint32)
    assert (sq.symbols_of_tokens(bad) == 0xFFFFFFFF).all()


def _codec_variant(tmp_path_factory, name, *flags):
    """The codec alone (no GPU half) built with extra flags, as a ctypes library."""
    import subprocess
    d = tmp_path_factory.mktemp(name)
    stub = d / "stub.c"
    stub.write_text('#include "sqz_gpu.h"\n#include <errno.h>\n'
                    'int sqz_gpu_stream_open(sqz_gpu_stream** s, int dev, const uint8_t* p, size_t n, uint32_t w, '
                    'uint32_t a, uint32_t b, uint32_t c, size_t k, uint32_t m) { (void)s; (void)dev; (void)p; (void)n; (void)w; (void)a; '
                    '(void)b; (void)c; (void)k; (void)m; return ENODEV; }\n'
                    'int sqz_gpu_stream_next(sqz_gpu_stream* s, const uint32_t** t, size_t* c) { (void)s; (void)t; (void)c; return ENODEV; }\n'
                    'void sqz_gpu_stream_close(sqz_gpu_stream* s) { (void)s; }\n'
                    'int sqz_gpu_expand_tokens(const uint32_t* t, size_t n, uint8_t* o, size_t b) { (void)t; (void)n; (void)o; (void)b; return ENODEV; }\n')
    so = d / ("lib%s.so" % name)
    subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-pthread", *flags,
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "sqz_b200", "csrc", "sqz_codec.c"),
                           str(stub), "-o", str(so)])
    L = C.CDLL(str(so))
    for fn_name in ("sqz_write_header", "sqz_init", "sqz_encode_tokens", "sqz_encode_symbols", "sqz_decompress_buffer"):
        fn = getattr(L, fn_name)
        fn.restype, fn.argtypes = _lib.SYMBOLS[fn_name]
    return L


@pytest.fixture(scope="module")
def selfcheck_lib(tmp_path_factory):
    """Built with -DSQZ_SELFCHECK: after every coded symbol it verifies that every leaf's cached plan
    equals a fresh one and that the tree is in the state the quick walk assumes."""
    return _codec_variant(tmp_path_factory, "sqzcheck", "-DSQZ_SELFCHECK")


@pytest.fixture(scope="module")
def selfcheck_small_blocks_lib(tmp_path_factory):
    """The self-check build with blocks of 16 x 16 tokens: spans, cuts and the one-by-one rests all happen
    within a few thousand tokens."""
    return _codec_variant(tmp_path_factory, "sqzchecksmall", "-DSQZ_SELFCHECK", "-DSQZ_PART_TOKENS=16")


@pytest.fixture(scope="module")
def tiny_log_lib(tmp_path_factory):
    """Built with a 64-entry change log: the two-thread coder's hand-off runs full all the time."""
    return _codec_variant(tmp_path_factory, "sqztinylog", "-DSQZ_LOG_SIZE=64")


def _skewed(n, seed):
    """Symbols with geometric probabilities: a deep, lopsided tree (plans longer than the usual length,
    leaves deeper than a plan holds) and, through the periodic shuffles, many reorderings."""
    rng = np.random.default_rng(seed)
    out = np.minimum(rng.geometric(0.35, n) - 1, 40).astype(np.uint32)
    for k in range(0, n, 4000):
        out[k:k + 4000] = (out[k:k + 4000] + (k // 4000) * 7) % 41
    lens = rng.integers(3, 258, n).astype(np.uint32)
    dist = np.minimum(2 ** rng.integers(0, 15, n) + rng.integers(0, 50, n), 32767).astype(np.uint32)
    is_match = rng.random(n) < 0.2
    is_match[:33000] = False                 # every distance below points at bytes that exist
    return np.where(is_match, lens << 16 | dist, out).astype(np.uint32)


@pytest.mark.parametrize("case", ["hello", "laozi.txt", "confucius.txt", "x64.elf", "mandrill.bmp", "skewed", "uniform"])
def test_quick_walk_plans_stay_valid(case, selfcheck_lib, inputs, oracle, reference):
    if case == "skewed":
        t = _skewed(60000, 5)
        nbytes = int(np.where(t >> 16 != 0, t >> 16, 1).sum())
    elif case == "uniform":
        t, nbytes = np.random.default_rng(3).integers(0, 256, 150000).astype(np.uint32), 150000
    else:
        d = inputs[case][:120000]
        t, nbytes = oracle_tokens(oracle, d, 15), d.size
    ours = sq.encode_tokens(t, nbytes, 15, lib=selfcheck_lib)          # aborts the process on a stale plan
    assert ours == sq.encode_tokens(t, nbytes, 15)
    assert ours == reference.encode_tokens(t, nbytes, 15)
    # the decoder keeps the same plans and, besides, a look-ahead table per tree (checked the same way)
    c = np.frombuffer(ours, np.uint8)
    out = np.zeros(nbytes, np.uint8)
    got = C.c_uint64()
    rc = selfcheck_lib.sqz_decompress_buffer(c.ctypes.data_as(_lib.u8p), c.size, out.ctypes.data_as(_lib.u8p), out.size,
                                             C.byref(got))
    assert rc == 0 and got.value == nbytes
    assert out.tobytes() == sq.decompress(ours) == reference.decompress(ours)


@pytest.mark.parametrize("name", ["hello", "abc40", "zeros4096", "empty", "one", "laozi.txt", "x64.elf"])
def test_decode_tokens_returns_what_was_encoded(name, inputs, oracle):
    """sqz_decode_tokens: the serial half of the decoder alone (the copy phase is sqz_gpu_expand_tokens)."""
    d = inputs[name]
    t = oracle_tokens(oracle, d, 15)
    got = sq.decode_tokens(sq.encode_tokens(t, d.size, 15))
    assert got.size == t.size and (got == t).all()


def test_long_mixed_stream_equals_the_reference_encoder(oracle, reference):
    """2.4 M tokens over text, ELF and image data (6 MiB of the bench corpus across its file seams): the
    quick, lazy and exact walks of the coder all take part; bytes must equal the unmodified reference's
    encoder on the same tokens, and both decoders must return the input."""
    from sqz_b200 import corpus
    d = corpus.synthetic(6 << 20, 3276897 - (1 << 20))
    t = oracle_tokens(oracle, d, 15)
    ours = sq.encode_tokens(t, d.size, 15)
    assert ours == reference.encode_tokens(t, d.size, 15)
    assert sq.encode_symbols(sq.symbols_of_tokens(t), d.size, 15) == ours
    assert sq.decompress(ours) == d.tobytes()
    got = sq.decode_tokens(ours)
    assert got.size == t.size and (got == t).all()


def _random_stream(rng):
    """A decodable token stream with a random alphabet, distribution and burstiness."""
    n = int(rng.integers(1, 30000))
    alphabet = rng.permutation(256)[: int(rng.integers(1, 257))]
    kind = int(rng.integers(0, 4))
    if kind == 0:        # uniform over the alphabet: constant near-ties, reorderings all the time
        lit = alphabet[rng.integers(0, alphabet.size, n)]
    elif kind == 1:      # zipf: a deep, lopsided tree
        lit = alphabet[np.minimum(rng.zipf(1.3, n) - 1, alphabet.size - 1)]
    elif kind == 2:      # phases: the distribution changes every few thousand tokens
        lit = np.concatenate([alphabet[(rng.integers(0, max(1, alphabet.size // 4), 2500) + off) % alphabet.size]
                              for off in rng.integers(0, alphabet.size, n // 2500 + 1)])[:n]
    else:                # two symbols taking turns: ties at the top of the tree
        lit = alphabet[np.arange(n) % min(2, alphabet.size)]
    p_match = float(rng.choice([0.0, 0.05, 0.5]))
    want_match = rng.random(n) < p_match
    lens = rng.integers(3, 258, n)
    dists = np.minimum(2 ** rng.integers(0, 15, n) + rng.integers(0, 9, n), 32767)
    toks, made = [], 0
    for k in range(n):
        if want_match[k] and made >= 1:
            toks.append(int(lens[k]) << 16 | int(min(dists[k], made)))
            made += int(lens[k])
        else:
            toks.append(int(lit[k]))
            made += 1
    return np.array(toks, np.uint32), made


@pytest.mark.parametrize("seed", range(40))
def test_random_streams_equal_the_reference_encoder(seed, reference):
    toks, nbytes = _random_stream(np.random.default_rng(1000 + seed))
    ours = sq.encode_tokens(toks, nbytes, 15)
    assert ours == reference.encode_tokens(toks, nbytes, 15)
    assert sq.encode_symbols(sq.symbols_of_tokens(toks), nbytes, 15) == ours
    got = sq.decode_tokens(ours)
    assert got.size == toks.size and (got == toks).all()
    assert sq.decompress(ours) == reference.decompress(ours)


@pytest.mark.parametrize("seed", range(12))
def test_random_streams_keep_the_self_check_quiet(seed, selfcheck_lib, reference):
    toks, nbytes = _random_stream(np.random.default_rng(2000 + seed))
    ours = sq.encode_tokens(toks, nbytes, 15, lib=selfcheck_lib)       # aborts on a stale plan or table
    assert ours == reference.encode_tokens(toks, nbytes, 15)
    c = np.frombuffer(ours, np.uint8)
    out = np.zeros(max(nbytes, 1), np.uint8)
    got = C.c_uint64()
    assert selfcheck_lib.sqz_decompress_buffer(c.ctypes.data_as(_lib.u8p), c.size, out.ctypes.data_as(_lib.u8p),
                                               out.size, C.byref(got)) == 0
    assert out[:nbytes].tobytes() == reference.decompress(ours)


@pytest.mark.parametrize("threads", [2, 4])
def test_model_thread_under_the_self_check(threads, selfcheck_lib, oracle, reference, inputs):
    """The model thread of the two-thread coder and of a crew in the self-check build: every span it
    accepts is replayed one by one from the same start (same weights, no reordering), every plan and
    weight is checked after every one-by-one symbol; bytes against the unmodified reference's encoder."""
    d = np.concatenate([inputs["confucius.txt"], inputs["x64.elf"][:150000], inputs["mandrill.bmp"][:60000]])
    t = oracle_tokens(oracle, d, 15)
    words = sq.symbols_of_tokens(t)
    assert sq.encode_symbols(words, d.size, 15, threads=threads, lib=selfcheck_lib) == reference.encode_tokens(t, d.size, 15)


@pytest.mark.parametrize("threads", [1, 2, 3])
@pytest.mark.parametrize("case", ["confucius.txt", "x64.elf", "mandrill.bmp", "skewed", "random"])
def test_spans_replayed_one_by_one(case, threads, selfcheck_small_blocks_lib, oracle, reference, inputs):
    """Blocks of 256 tokens in the self-check build: thousands of spans are accepted -- by the
    end-against-start test or by the walk over the culprits' rows --, each is replayed one by one from the
    same start and has to leave the same weights without a reordering; cuts, full-walk spans and the
    one-by-one rests in between; the bytes are the reference's."""
    import ctypes as C
    L = selfcheck_small_blocks_lib
    if case == "skewed":
        t = _skewed(60000, 9)
        nbytes = int(np.where(t >> 16 != 0, t >> 16, 1).sum())
    elif case == "random":
        t, nbytes = _random_stream(np.random.default_rng(77))
    else:
        d = inputs[case][:100000]
        t, nbytes = oracle_tokens(oracle, d, 15), d.size
    spans = C.c_uint64.in_dll(L, "sqz_selfcheck_spans")
    before = spans.value
    got = sq.encode_symbols(sq.symbols_of_tokens(t), nbytes, 15, threads=threads, lib=L)
    assert got == reference.encode_tokens(t, nbytes, 15)
    if t.size > 20000:
        assert spans.value - before > t.size // 2000, (spans.value - before, t.size)


@pytest.mark.parametrize("seed", range(30))
def test_two_thread_coder_gives_the_same_bytes(seed, reference):
    """coder_threads = 2: the model runs ahead on its own thread and logs code changes, the caller's
    thread packs the bits.  Forced here on streams of every size (the automatic choice starts at 64 Ki
    tokens); callback sinks too."""
    toks, nbytes = _random_stream(np.random.default_rng(3000 + seed))
    words = sq.symbols_of_tokens(toks)
    one = sq.encode_symbols(words, nbytes, 15, threads=1)
    two = sq.encode_symbols(words, nbytes, 15, threads=2)
    assert one == two == reference.encode_tokens(toks, nbytes, 15)
    # coder_threads >= 3: the model cuts the stream into segments, coder_threads - 1 emitters work on them
    assert sq.encode_symbols(words, nbytes, 15, threads=3 + seed % 4) == one
    if seed % 5 == 0:
        by_callback = sq.encode_symbols(words, nbytes, 15, file_mode=True, threads=1)
        assert sq.encode_symbols(words, nbytes, 15, file_mode=True, threads=2) == by_callback
        assert sq.encode_symbols(words, nbytes, 15, file_mode=True, threads=4) == by_callback


@pytest.mark.parametrize("seed", range(12))
def test_two_thread_coder_with_a_full_log(seed, tiny_log_lib, reference):
    """A 64-entry log is full after one reordering: the model has to wait for room, the emitter has to
    set changes aside that are not due yet -- and neither may wait for the other forever."""
    toks, nbytes = _random_stream(np.random.default_rng(4000 + seed))
    words = sq.symbols_of_tokens(toks)
    assert sq.encode_symbols(words, nbytes, 15, threads=2, lib=tiny_log_lib) == reference.encode_tokens(toks, nbytes, 15)


@pytest.mark.parametrize("threads", [1, 2, 3, 5])
@pytest.mark.parametrize("chunk", [1, 255, 256, 257, 5000, 100000])
def test_chunked_hand_over_like_sqz_compress(chunk, threads, oracle, reference, inputs):
    """sqz_compress feeds the coder chunk by chunk from the GPU stream; here the same from a host array."""
    d = inputs["arm64.elf"][:300000]
    t = oracle_tokens(oracle, d, 15)
    words = sq.symbols_of_tokens(t)
    L = _lib.load()
    out = np.zeros(d.size * 2 + 4096, np.uint8)
    bs = _bs(out)
    L.sqz_write_header(C.byref(bs), d.size, 15)
    s = _lib.State()
    L.sqz_init(C.byref(s))
    s.coder_threads = threads
    L.sqz_encode_symbols_chunked(C.byref(s), C.byref(bs), words.ctypes.data_as(_lib.u32p), words.size, chunk)
    assert s.error == 0 and s.tokens == t.size
    assert out[: bs.bytes].tobytes() == reference.encode_tokens(t, d.size, 15)


def test_two_thread_coder_errors():
    """A full sink and a word that is no symbol word stop both threads cleanly."""
    words = sq.symbols_of_tokens(np.random.default_rng(5).integers(0, 256, 200000).astype(np.uint32))
    L = _lib.load()
    buf = np.zeros(4096, np.uint8)
    bs = _bs(buf)
    L.sqz_write_header(C.byref(bs), 200000, 15)
    s = _lib.State()
    L.sqz_init(C.byref(s))
    for threads in (2, 4):
        bs = _bs(buf)
        L.sqz_write_header(C.byref(bs), 200000, 15)
        L.sqz_init(C.byref(s))
        s.coder_threads = threads
        L.sqz_encode_symbols(C.byref(s), C.byref(bs), words.ctypes.data_as(_lib.u32p), words.size)
        assert s.error == errno.E2BIG
    for where, what in ((150000, 256), (0, 285), (199999, 300), (77777, 284 | 31 << 9 | 3 << 14), (5, 257 | 30 << 14)):
        bad = words.copy()
        bad[where] = what
        for threads in (1, 2, 4):
            with pytest.raises(sq.SqzError) as e:
                sq.encode_symbols(bad, 200000, 15, threads=threads)
            assert e.value.errno == errno.EINVAL


def test_coders_agree_on_a_long_stream(oracle, reference):
    """3 MiB of the bench corpus (1.9 M tokens, 115 segments of the several-thread coder, some 500
    reorderings): one thread, two threads and crews of 2, 3 and 7 emitters give the reference's bytes."""
    from sqz_b200 import corpus
    d = corpus.synthetic(3 << 20, 5 << 20)
    t = oracle_tokens(oracle, d, 15)
    words = sq.symbols_of_tokens(t)
    want = reference.encode_tokens(t, d.size, 15)
    for threads in (1, 2, 3, 4, 8):
        assert sq.encode_symbols(words, d.size, 15, threads=threads) == want, threads


def test_decompress_into_the_callers_buffer(oracle, inputs):
    d = inputs["confucius.txt"]
    comp = sq.encode_tokens(oracle_tokens(oracle, d, 15), d.size, 15)
    buf = np.zeros(d.size + 100, np.uint8)
    out = sq.decompress(comp, into=buf)
    assert np.shares_memory(out, buf) and out.size == d.size and (out == d).all()
    with pytest.raises(ValueError):
        sq.decompress(comp, into=np.zeros(d.size - 1, np.uint8))


def test_header_bytes():
    """LSB-first fields in an MSB-first register, big-endian words."""
    c = sq.encode_tokens(np.zeros(0, np.uint32), 4096, 15)
    assert c[:9].hex() == "0008000000000000f0"
    assert sq.encode_tokens(np.zeros(0, np.uint32), 4096, 10)[8] == 0x50


def _bs(buf):
    bs = _lib.Bitstream()
    bs.data = buf.ctypes.data_as(_lib.u8p)
    bs.capacity = buf.size
    return bs


def test_header_rejects_bad_window():
    L = _lib.load()
    buf = np.zeros(64, np.uint8)
    for wb in (9, 16, 0, 255):
        bs = _bs(buf)
        L.sqz_write_header(C.byref(bs), 10, wb)
        assert bs.error == errno.EINVAL


def test_output_capacity_is_a_sticky_e2big():
    L = _lib.load()
    buf = np.zeros(16, np.uint8)            # room for the header only
    bs = _bs(buf)
    L.sqz_write_header(C.byref(bs), 1000, 15)
    assert bs.error == 0
    s = _lib.State()
    L.sqz_init(C.byref(s))
    t = np.arange(200, dtype=np.uint32) % 251
    L.sqz_encode_tokens(C.byref(s), C.byref(bs), t.ctypes.data_as(_lib.u32p), t.size)
    assert s.error == errno.E2BIG and bs.error == errno.E2BIG


@pytest.mark.parametrize("bad", [(258 << 16) | 1, (3 << 16) | 0x8000, (2 << 16) | 5, (3 << 16) | 0])
def test_tokens_the_decoder_would_reject_are_einval(bad):
    with pytest.raises(sq.SqzError) as e:
        sq.encode_tokens(np.array([65, 66, 67, bad], np.uint32), 300, 15)
    assert e.value.errno == errno.EINVAL


# 284 | 31 << 9 is bucket 27 with extra bits 31 = length 258, which the decoders reject (squeeze.h:529-545)
@pytest.mark.parametrize("bad", [256, 285, 511, 257 | 30 << 14, 284 | 31 << 9, 0xFFFFFFFF])
def test_words_that_are_not_symbol_words_are_einval(bad):
    with pytest.raises(sq.SqzError) as e:
        sq.encode_symbols(np.array([65, 66, 67, bad], np.uint32), 300, 15)
    assert e.value.errno == errno.EINVAL


def test_corrupt_streams_fail_cleanly(inputs, oracle):
    d = inputs["laozi.txt"]
    comp = bytearray(sq.encode_tokens(oracle_tokens(oracle, d, 15), d.size, 15))
    rng = np.random.default_rng(1)
    bad = 0
    for _ in range(40):
        c = bytearray(comp)
        for k in rng.integers(9, len(c), 3):
            c[k] ^= 1 << int(rng.integers(0, 8))
        try:
            out = sq.decompress(bytes(c))
            assert len(out) == d.size       # a flip may still decode to something of the right size
        except sq.SqzError as e:
            assert e.errno in (errno.EINVAL, errno.E2BIG)
            bad += 1
    assert bad > 0
    with pytest.raises(sq.SqzError):        # truncated
        sq.decompress(bytes(comp[: len(comp) // 2 // 8 * 8]))
    with pytest.raises(sq.SqzError):        # window byte out of range
        c = bytearray(comp); c[8] = 0xFF
        sq.decompress(bytes(c))


def test_compress_without_a_device_is_enodev():
    if sq.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(sq.SqzError) as e:
        sq.compress(b"no cpu fallback for the search", 15)
    assert e.value.errno == errno.ENODEV
    with pytest.raises(sq.SqzError) as e:
        sq.match_table(b"abcabcabc")
    assert e.value.errno == errno.ENODEV


@pytest.mark.parametrize("names", [("hello", "abc40"), ("laozi.txt", "hello"), ("zeros4096", "laozi.txt"), ("lorem3", "lorem3")])
def test_bitstream_position_after_decode_is_the_references(names, inputs, oracle):
    """After sqz_decompress the bitstream stands where the reference's word-at-a-time reader would
    (bitstream.h:65-95): `read` at the end of the stream's last word, whatever the decoder read ahead.
    Two streams stored back to back decode one after the other from the same sqz_bitstream."""
    L = _lib.load()
    parts = [inputs[n] for n in names]
    comps = [sq.encode_tokens(oracle_tokens(oracle, d, 15), d.size, 15) for d in parts]
    buf = np.frombuffer(b"".join(comps) + bytes(64), np.uint8).copy()
    bs = _bs(buf)
    bs.bytes = buf.size
    for d, comp in zip(parts, comps):
        start = int(bs.read)
        n, wb = C.c_uint64(), C.c_uint8()
        L.sqz_read_header(C.byref(bs), C.byref(n), C.byref(wb))
        assert bs.error == 0 and n.value == d.size and wb.value == 15
        out = np.zeros(max(d.size, 1), np.uint8)
        s = _lib.State()
        L.sqz_init(C.byref(s))
        L.sqz_decompress(C.byref(s), C.byref(bs), out.ctypes.data_as(_lib.u8p), d.size)
        assert s.error == 0 and out[: d.size].tobytes() == d.tobytes()
        assert int(bs.read) == start + len(comp), (names, int(bs.read), start, len(comp))
        assert 0 <= bs.bits < 64
        bs.bits = 0                     # the padding of the last word belongs to the stream just read
        bs.b64 = 0
