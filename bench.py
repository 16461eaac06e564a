#!/usr/bin/env python
"""bench.py -- match-search throughput of sqz-b200 on B200, beside the reference CPU codec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size BYTES]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4]): the synthetic corpus
-- the reference's six test/ files concatenated and repeated with seeded mutations -- cut
into contiguous shards, one per GPU, each with a max_dist look-back halo and a max_len
look-ahead halo.  Every rank holds --size bytes (1 GiB by default), so the job is
N x 1 GiB ("weak" scaling) and no data-path collective exists; the only exchange is one
integer per seam (the parse entry offset), handed over through torch.distributed.

A step = one pass of the hot path over the rank's shard: match table for every position
(squeeze.h:338-358 at every i) + greedy parse to the token stream (squeeze.h:377-394),
inputs and outputs resident in HBM.  `value` = input MB/s (1 MB = 1e6 B) of the whole
job, timed with CUDA events on the launching stream, max over ranks.  `e2e` = the same
pass through the host-buffer C-ABI sqz_gpu_match_table() from pinned host memory,
host<->device copies inside the timed region.

--impl reference times the UNMODIFIED reference codec (oracle/_ref, squeeze.compress)
on the host cores, on bounded samples of the same stream.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOW = 1 << 15
MIN_LEN, MAX_LEN, MAX_DIST = 3, 257, WINDOW - 1      # reference G1 rules, squeeze.h:13-15,342
SMEM_BYTES_PER_CLK_PER_SM = 128
SMS = 148
METRIC = "match_search_input_MBps"
TRAFFIC_1GIB = 1215603712 + 4293076480   # bytes, ncu, dominant kernel, one launch on a 1 GiB shard
KERNEL_ALU_INSTR = 354       # LOP3 + SHF per iteration of the hot loop (ncu source page, round 1 final kernel)
KERNEL_CC_PER_STEP = 16256   # 127 owned blocks x 32 positions x 4 distances per warp and iteration


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def cc_count(g0: int, n: int, md: int) -> int:
    """Candidate-compares for positions [g0, g0+n): sum of min(i, max_dist)  (SURVEY 8d)."""
    a, b = g0, g0 + n
    k = min(max(md, a), b)              # positions below k have i < md
    tri = (k - 1) * k // 2 - (a - 1) * a // 2 if k > a else 0
    return tri + (b - k) * md


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.path = tempfile.mktemp(prefix="sqz_clocks_", suffix=".csv")
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for k, nm in enumerate(names):
                    if f[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            busy = [c for c in sm if c > 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(power) if power else None)
        return out


# ----------------------------------------------------------------------------- reference arm
def reference_arm(args) -> None:
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import Oracle, Reference
    from sqz_b200 import corpus
    # oracle/_ref (the unmodified reference) when it was built; else the oracle's C restatement
    # of the same search + parse (kind "port"; it lacks the entropy stage, < 1 % of the time)
    kind = "reference" if Reference.available() else "port"
    cores = os.cpu_count() or 1
    slice_bytes = 128 << 10
    n_slices = 2 * cores
    total = args.size * args.gpus
    stride = max(total // n_slices, 1)
    slices = [corpus.synthetic(slice_bytes, (k * stride) // 4096 * 4096) for k in range(n_slices)]
    if kind == "port":
        Oracle.get().set_threads(1)

    def one_step() -> float:
        todo = list(range(n_slices))
        lock = threading.Lock()

        def worker():
            from oracle import Reference as R
            r = R(release=True) if kind == "reference" else None   # own handle; the C call releases the GIL
            while True:
                with lock:
                    if not todo:
                        return
                    k = todo.pop()
                if r is not None:
                    r.compress(slices[k], 15)
                else:
                    Oracle.get().tokens(slices[k], WINDOW)

        th = [threading.Thread(target=worker) for _ in range(cores)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        pass                                # a CPU loop has nothing to warm that matters; keep W for the record
    times = [one_step() for _ in range(args.steps)]
    sec = sum(times) / len(times)
    mbps = n_slices * slice_bytes / 1e6 / sec
    sample = ("%d slices x %d KiB of the synthetic stream, evenly spaced over %d GiB, one %s "
              "(window 2^15) per slice, work queue over %d threads; slices start with an empty window, "
              "which favours the reference by ~12%%"
              % (n_slices, slice_bytes >> 10, max(total >> 30, 1),
                 "squeeze.compress" if kind == "reference" else "oracle search+parse", cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": mbps, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": mbps, "unit": "MB/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mbps, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args) -> dict:
    return {
        "workload": "synthetic corpus (6 reference test/ files repeated + seeded mutations, SURVEY 8d config 5), "
                    "%d MiB per GPU, window 2^15, min_len 3, max_len 257, max_dist 32767" % (args.size >> 20),
        "bytes_per_gpu": args.size, "window": WINDOW, "min_len": MIN_LEN, "max_len": MAX_LEN,
        "max_dist": MAX_DIST, "parallelism": "shard%d" % args.gpus,
        "cache": "inputs (>= 1 GiB per GPU) are larger than the 126 MB L2; no explicit flush",
    }


# ----------------------------------------------------------------------------- our arm
def ours(args) -> None:
    import torch
    import torch.distributed as dist

    from sqz_b200 import _lib, corpus
    L = _lib.load()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run" % (args.gpus, world))
    if not torch.cuda.is_available() or L.sqz_gpu_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: sqz_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL has to say (its version banner when the
        # box sets NCCL_DEBUG) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    n = args.size
    g0 = rank * n                               # global offset of this rank's shard
    total = world * n
    back = min(g0, MAX_DIST)
    ahead = min(total - (g0 + n), MAX_LEN)
    host = corpus.synthetic(back + n + ahead, g0 - back)

    def check(rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %d %s" % (what, rc, L.sqz_gpu_last_error()))

    # pinned host buffers for the end-to-end leg
    def pinned(nbytes, dtype):
        p = L.sqz_gpu_host_alloc(nbytes)
        if not p:
            raise MemoryError("sqz_gpu_host_alloc(%d)" % nbytes)
        buf = (C.c_uint8 * nbytes).from_address(p)
        return p, np.frombuffer(buf, dtype=dtype)

    p_in, h_in = pinned(back + n + ahead, np.uint8)
    h_in[:] = host
    p_len, h_len = pinned(2 * (back + n + ahead), np.uint16)
    p_dist, h_dist = pinned(2 * (back + n + ahead), np.uint16)

    d_data = torch.empty(back + n + ahead + 64, dtype=torch.uint8, device=dev)
    d_data[: back + n + ahead].copy_(torch.from_numpy(host))
    d_table = torch.empty(n, dtype=torch.int32, device=dev)
    d_tokens = torch.empty(n, dtype=torch.int32, device=dev)
    d_work = torch.empty(L.sqz_gpu_parse_workspace(n), dtype=torch.uint8, device=dev)
    d_result = torch.zeros(2, dtype=torch.int64, device=dev)
    d_map = torch.zeros(512, dtype=torch.int16, device=dev)
    shard_ptr = d_data.data_ptr() + back
    stream = torch.cuda.current_stream()

    def step() -> None:
        s = stream.cuda_stream
        check(L.sqz_gpu_match_table_device(shard_ptr, back, n, ahead, MIN_LEN, MAX_LEN, MAX_DIST,
                                           d_table.data_ptr(), s), "match_table")
        entry = 0
        if world > 1:
            # seam hand-off: every shard publishes overshoot(entry) for all entries; a <= 8 step
            # chain on the host picks the real one.  One integer per seam, no payload collective.
            check(L.sqz_gpu_parse_exit_map_device(d_table.data_ptr(), n, MIN_LEN, MAX_LEN,
                                                  d_work.data_ptr(), d_map.data_ptr(), s), "exit_map")
            # 512 x u16 travel as 256 x i32 (NCCL has no 16-bit integer type)
            maps = [torch.empty(256, dtype=torch.int32, device=dev) for _ in range(world)]
            dist.all_gather(maps, d_map.view(torch.int32))
            hm = torch.stack(maps).cpu().numpy().view(np.uint16)
            for r in range(rank):
                entry = int(hm[r, entry])
        check(L.sqz_gpu_parse_device(shard_ptr, d_table.data_ptr(), n, entry, MIN_LEN, MAX_LEN,
                                     d_tokens.data_ptr(), n, d_work.data_ptr(), d_result.data_ptr(), s), "parse")

    def e2e_step() -> None:
        check(L.sqz_gpu_match_table(C.cast(p_in, _lib.u8p), back + n + ahead, WINDOW, MIN_LEN, MAX_LEN, MAX_DIST,
                                    C.cast(p_len, _lib.u16p), C.cast(p_dist, _lib.u16p)), "sqz_gpu_match_table")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg -------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    L.sqz_gpu_set_timing(1)
    L.sqz_gpu_match_kernel_seconds(1, None)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.sqz_gpu_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    sec = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    launches = L.sqz_gpu_launch_count() - launches0
    nl = C.c_uint64()
    t_match = L.sqz_gpu_match_kernel_seconds(1, C.byref(nl))
    L.sqz_gpu_set_timing(0)
    clocks = sampler.stop() if rank == 0 else {}
    n_tokens = int(d_result[0].item())

    # ---- end-to-end leg (host buffers through the C-ABI) -----------------------
    for _ in range(min(args.warmup, 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_sec = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)

    # spot check: the host-ABI table equals the device-resident one (same kernels, different plumbing)
    tab = d_table[: 1 << 20].cpu().numpy().view(np.uint32)
    assert ((tab >> 16) == h_len[back: back + (1 << 20)]).all() and \
           ((tab & 0xFFFF) == h_dist[back: back + (1 << 20)]).all(), "host ABI and device ABI disagree"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the unmodified reference on one core, bounded sample -----
    cpu = None
    try:
        from oracle import Reference
        ref = Reference.get(release=True)
        # 8 slices spread evenly over the shard (its head alone is text and flatters the reference)
        slices, each = 8, max(args.cpu_sample // 8, 4096)
        secs = 0.0
        for k in range(slices):
            off = (g0 + k * (n // slices)) // 4096 * 4096
            ref.compress(corpus.synthetic(each, off), 15)
            secs += ref.last_seconds
        sample_bytes = slices * each
        cpu = {"value": sample_bytes / 1e6 / secs, "unit": "MB/s", "cores": 1, "kind": "reference",
               "sample": "%d slices x %d KiB spread evenly over the shard, each through the unmodified reference's "
                         "squeeze.compress (oracle/_ref, -O3 -DNDEBUG, window 2^15, empty window at the slice start), "
                         "%.1f s on 1 of %d host cores" % (slices, each >> 10, secs, os.cpu_count() or 1)}
    except Exception as e:  # the oracle is only the yardstick; never let it sink the measurement
        cpu = {"value": None, "unit": "MB/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (e,)}

    # ---- whole codec: sqz_compress (GPU search + serial host entropy stage), bounded sample ----
    comp = None
    try:
        import sqz_b200 as sq
        sample = np.ascontiguousarray(host[back: back + min(n, args.compress_sample)])
        # one untimed call first, like the W warm-up steps of the main metric: it allocates the
        # pipeline's pinned and device buffers (two slots of 32 MiB chunks), which later calls reuse
        warm = sample[: min(sample.size, (33 << 20))]
        t0 = time.perf_counter()
        sq.compress(warm, 15)
        dt_warm = time.perf_counter() - t0
        st = {}
        t0 = time.perf_counter()
        blob = sq.compress(sample, 15, stats=st)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        back_again = sq.decompress(blob)
        dt_dec = time.perf_counter() - t0
        comp = {"value": sample.size / 1e6 / dt, "unit": "MB/s", "sample_bytes": int(sample.size),
                "compressed_bytes": len(blob), "seconds": dt, "search_wait_seconds": st["search_seconds"],
                "entropy_seconds": st["entropy_seconds"], "tokens": st["tokens"],
                "entropy_ns_per_token": st["entropy_seconds"] * 1e9 / max(st["tokens"], 1),
                "warmup": "one untimed sqz_compress of the first %d MiB (%.2f s: it allocates the pipeline's "
                          "pinned and device buffers)" % (warm.size >> 20, dt_warm),
                "decompress": {"value": sample.size / 1e6 / dt_dec, "unit": "MB/s", "seconds": dt_dec,
                               "round_trip_identical": back_again == sample.tobytes()},
                "host_threads": 2,
                "note": "sqz_compress(host in, host bitstream out): the serial adaptive-Huffman model bounds it "
                        "(SURVEY 7 H4) -- it runs on one host thread, a second one packs the bits; the search "
                        "runs ahead on the GPU and hands over symbol words (SURVEY 8f N3); sqz_decompress is "
                        "host only, one thread"}
    except Exception as e:
        comp = {"value": None, "error": repr(e)}

    ms_per_step = sec / args.steps * 1e3
    value = total / 1e6 / (sec / args.steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    f_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    smem_peak = SMEM_BYTES_PER_CLK_PER_SM * SMS * f_mhz * 1e6 / 1e9            # GB/s at the clock seen under load
    cc = cc_count(g0, n, MAX_DIST)
    achieved = cc * 4 / t_match / 1e9 if t_match > 0 else None
    # The bit-sliced kernel issues no load per candidate-compare, so the shared-memory figure can
    # exceed 1.  What binds it is the integer ALU pipe: 16 lanes/clk per SM sub-partition, i.e. half
    # a warp-instruction per clock.  Its fast path is KERNEL_ALU_INSTR ALU instructions per loop iteration
    # of KERNEL_CC_PER_STEP candidate-compares per warp (profiles/r01_match_table_ncu_full.txt).
    alu_ceiling = 0.5 * 4 * SMS * f_mhz * 1e6 * KERNEL_CC_PER_STEP / KERNEL_ALU_INSTR      # CC/s
    cc_per_s = cc / t_match if t_match > 0 else None
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_bytes = n * (1 + 4)                                                    # input read + table write
    line = {
        "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": total / 1e6 / e2e_sec, "unit": "MB/s",
                "h2d_bytes_per_step": int(back + n + ahead), "d2h_bytes_per_step": int(4 * (back + n + ahead)),
                "call": "sqz_gpu_match_table(host pinned in, host pinned len/dist out)", "steps": args.e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "smem", "kernel": "match_table",
            "achieved": achieved, "peak": smem_peak, "unit": "GB/s",
            "frac": achieved / smem_peak if achieved else None,
            "traffic": TRAFFIC_1GIB if n == (1 << 30) else None,
            "note": "algorithmic shared-memory bytes = 4 B per candidate-compare (SURVEY 8d), %d CC per launch; "
                    "peak = 128 B/clk/SM x 148 SMs x %.0f MHz (SM clock sampled under this load); "
                    "north_star fixes the smem compare bound, HBM is shown in 'hbm'" % (cc, f_mhz),
            "cc_per_launch": cc, "kernel_ms": t_match * 1e3, "kernel_launches_timed": int(nl.value),
            "kernel_share_of_step": (t_match * 1e3) / ms_per_step if ms_per_step else None,
            "binding_resource": {
                "name": "integer ALU pipe (LOP3/SHF), 0.5 warp-instr/clk per SM sub-partition",
                "alu_instr_per_warp_iteration": KERNEL_ALU_INSTR, "cc_per_warp_iteration": KERNEL_CC_PER_STEP,
                "ceiling_cc_per_s": alu_ceiling, "achieved_cc_per_s": cc_per_s,
                "frac": cc_per_s / alu_ceiling if cc_per_s else None,
                "note": "achieved counts the whole sqz_gpu_match_table_device call: bit-sliced kernel, edge tiles "
                        "and the finish kernel",
            },
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of v2::match_table<3,false> on a 1 GiB shard "
                            "(profiles/r01_dram_traffic_bench_size.csv: 1.22 GB read + 4.29 GB written; algorithmic "
                            "1.07 + 4.29 GB); the finish kernel adds 4.96 GB read + 0.82 GB written (it scans the table)",
        },
        "hbm": {"achieved": hbm_bytes / t_match / 1e9 if t_match > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_bytes / t_match / 1e9 / hbm_peak if t_match > 0 else None,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
        "cpu_baseline": cpu,
        "sqz_compress_e2e": comp,
        "tokens_per_shard": n_tokens,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries exactly one JSON line.  Libraries that print there (NCCL's version banner when
    # the box sets NCCL_DEBUG, torchrun children) are pointed at stderr for the whole run.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1 << 30, help="bytes per GPU")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=512 << 10)
    ap.add_argument("--compress-sample", type=int, default=256 << 20)
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
