"""The oracle (oracle/sqz_oracle.c) pinned against the unmodified reference (oracle/_ref),
the reference's second brute-force loop (bst.c) and the committed golden digests."""
import numpy as np
import pytest

from conftest import fnv

SMALL = ["zeros4096", "pat1234x1024", "hello", "abc40", "lorem3", "empty", "one", "two", "aaa", "aaaa", "laozi.txt"]
FILES = ["laozi.txt", "confucius.txt", "x64.elf", "arm64.elf", "mandrill.bmp", "mandrill.png"]

# complete known-answer vectors of the reference (window 2^15, memory mode)
KAT_HEX = {
    "zeros4096": "0008000000000000f0802717c0bcbdbe7efdfbf7efdfbf7efdfda00000000000",
    "pat1234x1024": "0008000000000000f0c0280e02409c5fc3bc5eefcfefef9f3e7cfdfbf7efd500",
    "hello": "c400000000000000f08934c9b1fd9047a809c39305c0d9ee8e4279062e000000",
}


def tok(lit=None, m=None):
    return lit if m is None else (m[0] << 16) | m[1]


def test_known_answer_tokens(oracle):
    """Token lists of SURVEY.md 8c (the reference's own test strings)."""
    z = oracle.tokens(bytes(4096), 1 << 15)
    assert list(z) == [0] + [tok(m=(257, 1))] * 15 + [tok(m=(240, 1))]
    p = oracle.tokens(bytes([1, 2, 3, 4]) * 1024, 1 << 15)
    assert list(p) == [1, 2, 3, 4] + [tok(m=(257, 4))] * 15 + [tok(m=(237, 4))]
    h = oracle.tokens(b"Hello World Hello.World Hello World", 1 << 15)
    assert list(h) == [ord(c) for c in "Hello World "] + [tok(m=(5, 12)), ord("."), tok(m=(11, 12)), tok(m=(6, 24))]
    a = oracle.tokens(b"abcabcdabcdeabcdefabcdefgabcdefabcdeabcd", 1 << 10)
    assert list(a) == [97, 98, 99, tok(m=(3, 3)), 100, tok(m=(4, 4)), 101, tok(m=(5, 5)), 102, tok(m=(6, 6)),
                       103, tok(m=(11, 13)), tok(m=(4, 5))]


@pytest.mark.parametrize("name", sorted(KAT_HEX))
def test_reference_known_answer_bytes(name, inputs, reference, golden):
    comp = reference.compress(inputs[name], 15)
    assert comp.hex() == KAT_HEX[name]
    assert fnv_hex(comp) == golden[name]["win"]["15"]["fnv_mem"]


def fnv_hex(b):
    from oracle import Oracle
    return "%016x" % Oracle.get().fnv(np.frombuffer(b, np.uint8))


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("wb", [10, 15])
def test_oracle_tokens_equal_reference_decisions(name, wb, inputs, oracle, reference, golden):
    """Every (len,pos)/literal decision the reference took == the oracle's."""
    d = inputs[name]
    comp = reference.compress(d, wb)
    assert fnv_hex(comp) == golden[name]["win"][str(wb)]["fnv_mem"]
    rt = reference.tokens(comp)
    ot = oracle.tokens(d, 1 << wb)
    assert rt.size == ot.size and (rt == ot).all()
    # and the reference's own encoder turns the oracle's tokens into the reference's bytes
    assert reference.encode_tokens(ot, d.size, wb) == comp
    assert reference.decompress(comp) == d.tobytes()


@pytest.mark.parametrize("name,start,size", [("confucius.txt", 0, 9000), ("confucius.txt", 12345, 9000),
                                             ("x64.elf", 1000, 7000), ("x64.elf", 903000, 7000),
                                             ("arm64.elf", 500001, 7000), ("mandrill.bmp", 77, 5000)])
def test_oracle_on_shifted_slices_of_the_fixtures(name, start, size, inputs, oracle, reference):
    """Different slice starts make the reference search different positions."""
    d = np.ascontiguousarray(inputs[name][start:start + size])
    for wb in (10, 12):
        rt = reference.tokens(reference.compress(d, wb))
        ot = oracle.tokens(d, 1 << wb)
        assert rt.size == ot.size and (rt == ot).all()


@pytest.mark.parametrize("name", SMALL + ["confucius.txt"])
@pytest.mark.parametrize("wb", [10, 15])
def test_brute_force_table_golden(name, wb, inputs, oracle, golden):
    d = inputs[name]
    g = golden[name]["win"][str(wb)]
    ln, ds = oracle.match_table(d, 1 << wb)
    assert fnv(oracle, ln) == g["fnv_len"] and fnv(oracle, ds) == g["fnv_dist"]
    t, end = oracle.tokens_from_table(d, ln, ds)
    assert end == d.size and t.size == g["tokens"] and fnv(oracle, t) == g["fnv_tokens"]
    assert int((t >> 16 != 0).sum()) == g["matches"]
    if "token_list" in g:
        assert [int(x) for x in t] == g["token_list"]


@pytest.mark.parametrize("name", FILES)
@pytest.mark.parametrize("wb", [10, 15])
def test_fast_oracle_golden(name, wb, inputs, oracle, golden):
    """Oracle B (hash chains) reproduces the brute-force digests on every fixture."""
    d = inputs[name]
    g = golden[name]["win"][str(wb)]
    ln, ds = oracle.match_table(d, 1 << wb, fast=True)
    assert fnv(oracle, ln) == g["fnv_len"] and fnv(oracle, ds) == g["fnv_dist"]
    t, end = oracle.tokens_from_table(d, ln, ds)
    assert end == d.size and fnv(oracle, t) == g["fnv_tokens"]


def _cases():
    rng = np.random.default_rng(7)
    yield "random", rng.integers(0, 256, 5000, dtype=np.uint8)
    yield "two_symbols", rng.integers(0, 2, 6000, dtype=np.uint8)
    yield "runs", np.repeat(rng.integers(0, 4, 60, dtype=np.uint8), rng.integers(1, 400, 60))
    yield "period7", np.tile(rng.integers(0, 256, 7, dtype=np.uint8), 700)
    yield "zeros_then_noise", np.concatenate([np.zeros(3000, np.uint8), rng.integers(0, 256, 2000, dtype=np.uint8)])


@pytest.mark.parametrize("case", [c[0] for c in _cases()])
@pytest.mark.parametrize("rules", [(3, 257, None), (2, 254, None), (2, 254, "window"), (4, 16, None), (3, 40, 100)])
def test_fast_oracle_equals_brute_force(case, rules, oracle):
    d = dict(_cases())[case]
    mn, mx, md = rules
    for window in (1 << 10, 1 << 12):
        kw = dict(min_len=mn, max_len=mx, max_dist=(window if md == "window" else md))
        a = oracle.match_table(d, window, **kw)
        b = oracle.match_table(d, window, fast=True, **kw)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


@pytest.mark.parametrize("text", [
    b"abcabcdabcdeabcdefabcdefgabcdefabcdeabcd",                 # bst.c:322 / test.c:547
    b"abcabcdabcdeabcdefabcdefgabcdefabcdeabcd" * 2,
    b"0123456789abcdef0123456789ABCDEF" * 4,
])
def test_rule_set_iii_against_bst_c(text, oracle, reference):
    """bst.c:230-252 lz77_find (min 2, max 254, dist <= window) == the parameterised oracle."""
    d = np.frombuffer(text, np.uint8)
    for window in (8, 16, 64, 1024, 65535):
        rl, rd = reference.bst_table(d, window)
        ol, od = oracle.match_table(d, 1 << 16, min_len=2, max_len=254, max_dist=window)
        assert (rl == ol).all() and (rd == od).all()


def test_rule_set_iii_on_a_fixture_slice(inputs, oracle, reference):
    d = np.ascontiguousarray(inputs["x64.elf"][904000:909000])
    rl, rd = reference.bst_table(d, 1024)
    ol, od = oracle.match_table(d, 1 << 16, min_len=2, max_len=254, max_dist=1024)
    assert (rl == ol).all() and (rd == od).all()


def test_table_then_walk_equals_interleaved_parse(inputs, oracle):
    """full table + walk over next[i] == the reference's interleaved search/parse."""
    d = inputs["laozi.txt"]
    for wb in (10, 15):
        ln, ds = oracle.match_table(d, 1 << wb)
        t, end = oracle.tokens_from_table(d, ln, ds)
        assert end == d.size and (t == oracle.tokens(d, 1 << wb)).all()
