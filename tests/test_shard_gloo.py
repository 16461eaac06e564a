"""The N>1 path on CPU: world_size-2 and -3 process groups over gloo.  Every rank owns one
contiguous shard (with halos), computes its match table, publishes its exit map, receives
the others' through an all-gather, chains them into its true parse entry and emits its
tokens; rank 0 concatenates.  The compute engine here is the oracle (there is no GPU in this
test): what is under test is the host logic -- shard plan, halos, seam hand-off, concatenation
-- which bench.py and a multi-GPU caller run unchanged around the CUDA entry points."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

MIN_LEN, MAX_LEN, WINDOW = 3, 257, 1 << 12


def cpu_exit_map(ln, n, min_len, max_len):
    """What sqz_gpu_parse_exit_map_device computes, restated for the CPU test."""
    out = np.zeros(512, dtype=np.uint16)
    for e in range(max_len):
        i = e
        while i < n:
            i += int(ln[i]) if ln[i] >= min_len else 1
        out[e] = i - n
    return out


def worker(rank, world, port, total, q):
    import sys
    sys.path.insert(0, ROOT)
    from oracle import Oracle
    from sqz_b200 import corpus, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle.get()
    o.set_threads(2)
    data = corpus.synthetic(total, 3276897 - 20000)            # straddles a repetition boundary
    s = shard.plan(total, world, WINDOW - 1, MAX_LEN)[rank]
    local = np.ascontiguousarray(data[s.lo:s.hi])              # shard + halos only
    # table of the owned positions, computed from the local bytes alone
    ln, ds = o.match_table(local, WINDOW, first=s.back, count=s.n + s.ahead, fast=True)
    # the look-ahead halo lets matches run past the shard end; cap at the global end is implicit
    ln, ds = ln[:s.n], ds[:s.n]
    # 512 x u16 travel as 256 x i32 (neither gloo nor NCCL moves 16-bit integers)
    mine = torch.from_numpy(cpu_exit_map(ln, s.n, MIN_LEN, MAX_LEN).view(np.int32).copy())
    maps = [torch.zeros(256, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(maps, mine)                                # the only exchange: 1 KiB per rank
    entries = shard.chain_entries([m.numpy().view(np.uint16) for m in maps])
    toks, end = o.tokens_from_table(local[s.back:s.back + s.n], ln, ds, MIN_LEN, start=entries[rank])
    assert end - s.n == entries[rank + 1]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(toks, gathered, dst=0)
    if rank == 0:
        whole = o.tokens(data, WINDOW)
        cat = np.concatenate(gathered)
        q.put((bool(cat.size == whole.size and (cat == whole).all()), int(whole.size), entries))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 60000), (3, 50001)])
def test_sharded_parse_equals_single_shard(world, total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, n_tokens, entries = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n_tokens > 0
    assert entries[0] == 0 and entries[-1] == 0                # a complete parse ends exactly at the end


def mailbox_worker(rank, world, port, total, steps, q):
    """The bench's N>1 step with the product's host logic (shard.plan, shard.Mailbox): seam exchange
    and token concatenation without a collective -- gloo only provides the rendezvous and the barrier,
    as NCCL does in bench.py.  The gather buffer is a shared-memory array standing in for the
    IPC-mapped device buffer; every rank writes its tokens at the offset the mailbox gives it."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import Oracle
    from sqz_b200 import corpus, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle.get()
    o.set_threads(2)
    path = "/dev/shm/sqz_test_mailbox_%d" % port
    mb = shard.Mailbox(path, rank, world, create=(rank == 0), timeout=60)
    if rank == 0:
        mb.put_blob(b"gather-handle")
    assert mb.get_blob()[:13] == b"gather-handle"
    data = corpus.synthetic(total, 3276897 - 20000)
    s = shard.plan(total, world, WINDOW - 1, MAX_LEN)[rank]
    local = np.ascontiguousarray(data[s.lo:s.hi])
    ln, ds = o.match_table(local, WINDOW, first=s.back, count=s.n + s.ahead, fast=True)
    ln, ds = ln[:s.n], ds[:s.n]
    gather_path = path + ".tokens"
    if rank == 0:
        np.zeros(total, np.uint32).tofile(gather_path)
    dist.barrier()
    gather = np.memmap(gather_path, dtype=np.uint32, mode="r+", shape=(total,))
    ok = True
    for step in range(1, steps + 1):
        entry, entries = mb.entry(step, cpu_exit_map(ln, s.n, MIN_LEN, MAX_LEN)[:MAX_LEN])
        toks, end = o.tokens_from_table(local[s.back:s.back + s.n], ln, ds, MIN_LEN, start=entry)
        at = mb.offset(step, toks.size)
        gather[at:at + toks.size] = toks
        gather.flush()
        mb.barrier(step)
        if rank == 0:
            counts = mb.counts(step)
            whole = o.tokens(data, WINDOW)
            got = np.array(gather[: sum(counts)])
            ok = ok and got.size == whole.size and bool((got == whole).all())
            ok = ok and shard.offsets(counts)[-1] + counts[-1] == whole.size
    if rank == 0:
        q.put(ok)
    dist.barrier()
    mb.close(unlink=(rank == 0))
    if rank == 0:
        os.unlink(gather_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 60000), (4, 70001)])
def test_mailbox_seams_and_concatenation(world, total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=mailbox_worker, args=(r, world, port, total, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_mailbox_times_out_instead_of_hanging(tmp_path):
    from sqz_b200 import shard
    mb = shard.Mailbox(str(tmp_path / "mb"), 1, 2, create=True, timeout=0.2)
    with pytest.raises(TimeoutError):
        mb.entry(1, np.zeros(257, np.uint16))        # rank 0 never publishes
    mb.close(unlink=True)
