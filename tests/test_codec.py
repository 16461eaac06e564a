"""Host entropy stage (csrc/sqz_codec.c) without a GPU: bit-exact against the golden digests
made with the unmodified reference, round trips, header bytes, error behaviour."""
import ctypes as C
import errno

import numpy as np
import pytest

import sqz_b200 as sq
from conftest import fnv
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
