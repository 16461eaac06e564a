"""Writes profiles/r02_roofline_inputs.json, the file bench.py takes its roofline inputs from.

  * instruction counts of the hot loop: from the SASS of the library as built (tools/hot_loop.py);
  * DRAM traffic per input byte: from an ncu pass over one bench step at full size,
        ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
            -k regex:"match_table|finish_marked|combine_slices" -c 12 --clock-control none --csv \
            --log-file gpurun_out/r02_traffic.csv python bench.py --steps 1 --warmup 1 --e2e-steps 1 ...
    (the last launch of each kernel in the CSV is the timed step's);
  * the commit both were taken at.

    python tools/roofline_inputs.py gpurun_out/r02_traffic.csv <input bytes of the shard>
"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from hot_loop import hot_loop

traffic_csv = sys.argv[1] if len(sys.argv) > 1 else None
shard_bytes = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 30
out = hot_loop()
out["source"] = ("instruction counts: cuobjdump -sass of sqz_b200/lib/libsqz_b200.so (tools/hot_loop.py); "
                 "DRAM bytes: ncu dram__bytes_read.sum + dram__bytes_write.sum of one bench step on a %d MiB shard "
                 "(profiles/r02_dram_traffic_bench_size.csv)" % (shard_bytes >> 20))
out["commit"] = subprocess.run(["git", "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
if traffic_csv and os.path.exists(traffic_csv):
    rows = list(csv.reader(l for l in open(traffic_csv) if l.startswith('"')))
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = {}
    for r in rows[1:]:
        per.setdefault(r[ii], {"kernel": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
    last = {}
    for _, d in sorted(per.items(), key=lambda kv: int(kv[0])):
        name = d["kernel"].split("(")[0].replace("void ", "")
        last[name] = d                                   # later launches overwrite: the timed step's remain
    kernels, total = {}, 0.0
    for name, d in last.items():
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        kernels[name] = {"dram_bytes_read": d.get("dram__bytes_read.sum"), "dram_bytes_write": d.get("dram__bytes_write.sum"),
                         "duration_ns": d.get("gpu__time_duration.sum")}
        total += b
    main = [v for k, v in kernels.items() if "match_table<3, 0>" in k or "match_table<3,0>" in k]
    out["dram_by_kernel"] = kernels
    out["dram_bytes_per_input_byte"] = total / shard_bytes
    if main:
        m = main[0]
        out["dram_bytes_per_input_byte_main_kernel"] = (m["dram_bytes_read"] + m["dram_bytes_write"]) / shard_bytes
    out["traffic_note"] = ("whole sqz_gpu_match_table_device call on a %d MiB shard: %.2f B of DRAM traffic per input byte "
                           "against 5 B algorithmic (1 B in + 4 B table out)" % (shard_bytes >> 20, total / shard_bytes))
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_roofline_inputs.json"), "w"), indent=1)
print(json.dumps({k: out[k] for k in out if k not in ("histogram", "dram_by_kernel")}, indent=1))
