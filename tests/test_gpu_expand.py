"""The decoder's copy phase on the GPU: sqz_gpu_expand_tokens == the reference's
byte-by-byte execution of the same tokens (squeeze.h:502-551)."""
import errno

import numpy as np
import pytest

import sqz_b200 as sq
from sqz_b200 import corpus

pytestmark = pytest.mark.gpu

NAMES = ["hello", "abc40", "zeros4096", "pat1234x1024", "lorem3", "one", "two", "aaa", "aaaa",
         "laozi.txt", "confucius.txt", "x64.elf", "arm64.elf", "mandrill.bmp", "mandrill.png"]


def tokens_of(oracle, d, wb=15):
    return oracle.tokens_from_table(d, *oracle.match_table(d, 1 << wb, fast=True))[0]


@pytest.mark.parametrize("name", NAMES)
def test_expand_reproduces_the_input(name, inputs, oracle):
    d = inputs[name]
    assert sq.expand_tokens(tokens_of(oracle, d), d.size) == d.tobytes()


@pytest.mark.parametrize("name", ["hello", "laozi.txt", "arm64.elf", "mandrill.bmp"])
def test_decompress_gpu_equals_host_and_reference(name, inputs, oracle, reference):
    d = inputs[name]
    comp = reference.compress(d, 15)
    st = {}
    assert sq.decompress_gpu(comp, stats=st) == d.tobytes() == sq.decompress(comp)
    assert st["tokens"] == tokens_of(oracle, d).size


def test_long_chains():
    """Runs: every byte of a (257,1) match hangs on the byte before it, so a 3 MiB run of one
    value is one chain of 3M hops -- 22 rounds of doubling; period-3 and period-1000 likewise."""
    for unit, total in ((b"\0", 3 << 20), (b"abc", 1 << 20), (bytes(range(250)) * 4, 2 << 20)):
        data = (unit * (total // len(unit) + 1))[:total]
        toks = list(unit)
        at = len(unit)
        while at < total:
            ln = min(257, total - at)
            if ln < 3:
                toks += list(data[at:at + ln])
            else:
                toks.append(ln << 16 | len(unit))
            at += ln
        assert sq.expand_tokens(np.array(toks, np.uint32), total) == data


def test_synthetic_16mib(oracle):
    d = corpus.synthetic(16 << 20, 3276897 * 3 - 4000000)
    t = sq.tokens(d)
    assert sq.expand_tokens(t, d.size) == d.tobytes()


def test_bad_token_streams_are_einval():
    lit = np.arange(10, dtype=np.uint32)
    for toks, n in ((np.r_[lit, (5 << 16) | 11].astype(np.uint32), 15),      # reaches before the start
                    (np.r_[lit, (5 << 16) | 0].astype(np.uint32), 15),       # distance 0
                    (lit, 11), (lit, 9),                                      # wrong announced size
                    (np.r_[(3 << 16) | 1, lit].astype(np.uint32), 13)):      # match as first token
        with pytest.raises(sq.SqzError) as e:
            sq.expand_tokens(toks, n)
        assert e.value.errno == errno.EINVAL
    assert sq.expand_tokens(np.zeros(0, np.uint32), 0) == b""
