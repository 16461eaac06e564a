"""Device-buffer entry points (include/sqz_gpu.h) on torch CUDA tensors.

torch is plumbing here: it owns the device memory and the stream; every kernel launched is
this library's own (libsqz_b200.so).  A "shard" is a view into a uint8 CUDA tensor that also
holds its halos: positions [first, first+n) of `buf`, with `back` readable bytes before and
`ahead` after.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .api import G1_MAX_LEN, G1_MIN_LEN, _check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need(t: torch.Tensor, what: str, dtype, numel: int, device) -> None:
    """The library writes through raw pointers: a wrong dtype, size or device corrupts memory."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == device and t.dtype == dtype
            and t.is_contiguous() and t.numel() >= numel):
        raise ValueError(f"{what} must be a contiguous {dtype} CUDA tensor on {device} with at least {numel} elements")


def match_table(buf: torch.Tensor, first: int, n: int, back: int, ahead: int, min_len: int = G1_MIN_LEN,
                max_len: int = G1_MAX_LEN, max_dist: int = (1 << 15) - 1, out: torch.Tensor | None = None) -> torch.Tensor:
    """Packed table words (len << 16 | dist) for buf[first : first+n]  (squeeze.h:340-358 at every i)."""
    _need(buf, "buf", torch.uint8, first + n + ahead, buf.device if isinstance(buf, torch.Tensor) else None)
    if first - back < 0 or n < 0:
        raise ValueError("the look-back halo reaches before the buffer")
    if out is None:
        out = torch.empty(max(n, 1), dtype=torch.int32, device=buf.device)
    else:
        _need(out, "out", torch.int32, n, buf.device)
    rc = _lib.load().sqz_gpu_match_table_device(buf.data_ptr() + first, back, n, ahead, min_len, max_len, max_dist,
                                                out.data_ptr(), _stream())
    _check(rc, "sqz_gpu_match_table_device")
    return out[:n]


def exit_map(table: torch.Tensor, n: int, min_len: int = G1_MIN_LEN, max_len: int = G1_MAX_LEN) -> torch.Tensor:
    """exit_map[e] = overshoot of the shard when entered at offset e (512 x int16 on the device)."""
    L = _lib.load()
    _need(table, "table", torch.int32, n, table.device if isinstance(table, torch.Tensor) else None)
    work = torch.empty(L.sqz_gpu_parse_workspace(n), dtype=torch.uint8, device=table.device)
    out = torch.zeros(512, dtype=torch.int16, device=table.device)
    rc = L.sqz_gpu_parse_exit_map_device(table.data_ptr(), n, min_len, max_len, work.data_ptr(), out.data_ptr(), _stream())
    _check(rc, "sqz_gpu_parse_exit_map_device")
    torch.cuda.current_stream().synchronize()      # `work` must outlive the kernels
    return out


def parse(buf: torch.Tensor, first: int, table: torch.Tensor, n: int, entry: int = 0, min_len: int = G1_MIN_LEN,
          max_len: int = G1_MAX_LEN, symbols: bool = False):
    """Greedy parse of one shard (squeeze.h:337,377-394): (tokens int32[count], overshoot).
    symbols=True: symbol words (include/sqz_gpu.h) instead of plain tokens."""
    L = _lib.load()
    dev = buf.device if isinstance(buf, torch.Tensor) else None
    _need(buf, "buf", torch.uint8, first + n, dev)
    _need(table, "table", torch.int32, n, dev)
    work = torch.empty(L.sqz_gpu_parse_workspace(n), dtype=torch.uint8, device=dev)
    tokens = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    result = torch.zeros(2, dtype=torch.int64, device=dev)
    fn = L.sqz_gpu_parse_symbols_device if symbols else L.sqz_gpu_parse_device
    rc = fn(buf.data_ptr() + first, table.data_ptr(), n, entry, min_len, max_len,
            tokens.data_ptr(), n, work.data_ptr(), result.data_ptr(), _stream())
    _check(rc, "sqz_gpu_parse_symbols_device" if symbols else "sqz_gpu_parse_device")
    count, overshoot = (int(x) for x in result.cpu())
    return tokens[:count], overshoot
