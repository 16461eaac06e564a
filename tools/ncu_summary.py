"""Text summary of one kernel of an ncu report (`ncu --set full --import-source on`): launch shape,
pipe utilisation, stall ratios, DRAM bytes, and the SASS regions of the source page grouped by how
often they ran (the hot loop is the region with the largest share of instructions).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep <kernel regex> <launch index among matches> > profiles/rNN_<kernel>_ncu_full.txt
"""
import collections, csv, io, re, subprocess, sys

rep, pattern, skip = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0
sel = ["--kernel-name", "regex:" + pattern, "--launch-skip", str(skip), "--launch-count", "1"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", *sel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = re.compile(r"^(Kernel Name|dram__bytes_(read|write)\.sum(\.per_second)?$|gpu__time_duration\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|"
                  r"launch__(block_size|grid_size|occupancy_limit_\w+|registers_per_thread|shared_mem_per_block_dynamic)$|sm__cycles_elapsed\.avg\.per_second|"
                  r"sm__inst_executed_pipe_(alu|fma|lsu|xu|uniform)\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__thread_inst_executed_per_inst_executed\.ratio|smsp__warps_active\.avg\.per_cycle_active|lts__t_sector_hit_rate\.pct)")
for h, u, v in sorted(zip(hdr, units, vals)):
    if keep.match(h):
        print("%s [%s] = %s" % (h, u, v))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *sel], capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = rows[0]
ci, ce, ct, cs = h.index("Source"), h.index("Instructions Executed"), h.index("Avg. Threads Executed"), h.index("# Samples")
seen, ins = set(), []
for r in rows[1:]:                       # the page lists the SASS once per view: keep the first
    if len(r) > cs and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0])
        ins.append((r[ci].strip(), float(r[ce]), float(r[ct]), float(r[cs])))
total_i = sum(x[1] for x in ins) or 1.0
total_s = sum(x[3] for x in ins) or 1.0
print("\n# source page, SASS regions by execution count (consecutive instructions executed equally often, within 2 %%); "
      "%d instructions, %.4g executed" % (len(ins), total_i))
regions, a = [], 0
for k in range(1, len(ins) + 1):
    if k == len(ins) or abs(ins[k][1] - ins[a][1]) > 0.02 * max(ins[a][1], 1.0):
        regions.append((a, k))
        a = k
for a, b in regions:
    n = b - a
    ex = sum(x[1] for x in ins[a:b])
    if ex / total_i < 0.003 and sum(x[3] for x in ins[a:b]) / total_s < 0.01:
        continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[0]).split()[0].split(".")[0] for x in ins[a:b])
    thr = sum(x[2] * x[1] for x in ins[a:b]) / max(ex, 1.0)
    print("sass[%4d..%4d) n=%3d exec/instr=%.3e share_of_instructions=%5.1f%% share_of_samples=%5.1f%% avg_threads=%4.1f  %s"
          % (a, b, n, ex / n, 100 * ex / total_i, 100 * sum(x[3] for x in ins[a:b]) / total_s, thr, ops.most_common(8)))
