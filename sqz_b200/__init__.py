"""sqz_b200 -- B200-native LZ77 match search for the sqz codec (leok7v/sqz).

Only the hot path of the reference lives here: the longest-match search and the
greedy parse on the GPU (csrc/sqz_gpu.cu), the host entropy stage that consumes
the token stream unchanged (csrc/sqz_codec.c), and this thin ctypes mirror of
the reference's codec interface.  See DESIGN.md.
"""
from .api import (SqzError, capacity, compress, decode_tokens, decompress, decompress_gpu, device_count, expand_tokens, encode_symbols, encode_tokens, launch_count,
                  match_table, read_header, release, select_kernel, symbols_of_tokens, tokens, tokens_multi)

__all__ = ["SqzError", "capacity", "compress", "decode_tokens", "decompress", "decompress_gpu", "device_count", "expand_tokens", "encode_symbols", "encode_tokens",
           "launch_count", "match_table", "read_header", "release", "select_kernel", "symbols_of_tokens", "tokens", "tokens_multi"]
