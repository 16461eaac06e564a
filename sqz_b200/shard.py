"""Multi-GPU sharding of the match search (host logic only, no compute).

The path shards by contiguous byte ranges: the match table of
shard g depends only on its own bytes plus a look-back halo of max_dist bytes and a
look-ahead halo of max_len bytes, so no data-path collective exists.  The greedy parse
has one scalar dependency per seam: the first parse position of shard g+1 is the
overshoot of shard g's last token.  Every shard publishes exit_map[e] = "overshoot I
produce when entered at offset e" (sqz_gpu_parse_exit_map_device); chaining those maps
gives every shard its true entry after one tiny all-gather.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    first: int      # global offset of the first owned position
    n: int          # owned positions
    back: int       # look-back halo bytes available before `first`
    ahead: int      # look-ahead halo bytes available after first + n

    @property
    def lo(self) -> int:
        """Global offset of the first byte the shard needs."""
        return self.first - self.back

    @property
    def hi(self) -> int:
        """One past the last byte the shard needs."""
        return self.first + self.n + self.ahead


def plan(total: int, world: int, max_dist: int, max_len: int) -> list[Shard]:
    """Cut [0, total) into `world` contiguous shards of (almost) equal size."""
    if world < 1:
        raise ValueError("world must be >= 1")
    base, extra = divmod(total, world)
    out, first = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(Shard(r, first, n, min(first, max_dist), min(total - (first + n), max_len)))
        first += n
    return out


def chain_entries(exit_maps: Sequence[np.ndarray], first_entry: int = 0) -> list[int]:
    """entry[0] = first_entry, entry[g+1] = exit_maps[g][entry[g]]; returns world+1 values
    (the last one is the overshoot past the end of the data, 0 for a complete parse)."""
    entries = [int(first_entry)]
    for m in exit_maps:
        entries.append(int(np.asarray(m).astype(np.int64)[entries[-1]] & 0xFFFF))
    return entries
