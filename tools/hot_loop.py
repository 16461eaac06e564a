"""Instruction counts of the match kernel's hot loop, read from the SASS of the built library.

The loop over the 32 bit offsets of a group of 128 distances (`for sh` in match_bitsliced.cuh) is
the smallest backward branch around the four SHFL.DOWN of the look-ahead exchange.  Its fast path
is the straight line from the loop head to the branch that skips the scalar path, plus the tail
from that branch's target to the backward branch.  Per iteration a warp covers 127 owned blocks x
32 positions x 4 distances = 16,256 candidate-compares.

    python tools/hot_loop.py [libsqz_b200.so] [--json]
"""
import json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass(lib, mangled_part):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    keep, on = [], False
    for line in out.splitlines():
        if "Function :" in line:
            on = mangled_part in line
        elif on:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                keep.append((int(m.group(1), 16), m.group(2).strip()))
    return keep


def hot_loop(lib=None, kernel="match_tableILi3ELb0ELi4E"):
    lib = lib or os.path.join(ROOT, "sqz_b200", "lib", "libsqz_b200.so")
    ins = sass(lib, kernel)
    shfl = [a for a, t in ins if t.startswith("SHFL.DOWN")]
    assert len(shfl) >= 4, "look-ahead exchange not found"
    first, last = shfl[0], shfl[-1]
    loops = []
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and a > last and int(m.group(1), 16) <= first:
            loops.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
    size, head, back = min(loops)
    def op(t):
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        return t.split()[0].split(".")[0]

    def fwd(after, upto):
        """first conditional forward branch at an address in (after, upto) that lands inside the loop"""
        for a, t in ins:
            m = re.search(r"^@!?U?P\d+\s+BRA\s+0x([0-9a-f]+)", t)
            if m and after < a < upto and a < int(m.group(1), 16) <= back:
                return a, int(m.group(1), 16)
        return None, None

    addr = [a for a, _ in ins]
    text = dict(ins)
    # adaptive gate: an if/else right after the exchange (the body with the need masks, the body without)
    g_at, g_to = fwd(last, back)
    bodies = None
    prev = addr[addr.index(g_to) - 1]
    m = re.match(r"^BRA\s+0x([0-9a-f]+)", text[prev])
    if m and int(m.group(1), 16) > g_to:
        join = int(m.group(1), 16)
        one = [t for a, t in ins if g_at < a <= prev]
        two = [t for a, t in ins if g_to <= a < join]
        bodies = sorted([one, two], key=len)            # [quiet body, body with the need masks]
        skip_at, skip_to = fwd(join - 1, back)
        common = [t for a, t in ins if head <= a <= g_at or join <= a <= skip_at or skip_to <= a <= back]
    else:                                               # one body: the branch found is the one over the scalar path
        skip_at, skip_to = g_at, g_to
        common = [t for a, t in ins if head <= a <= skip_at or skip_to <= a <= back]
        bodies = [[], []]
    assert skip_at is not None, "branch over the scalar path not found"

    def count(lines):
        hist = {}
        for t in lines:
            hist[op(t)] = hist.get(op(t), 0) + 1
        return hist
    # planes only the gated body compares: a conditional forward branch in front of the exchange that lands in front of it
    p_at = p_to = None
    for a, t in ins:
        m = re.search(r"^@!?U?P\d+\s+BRA\s+0x([0-9a-f]+)", t)
        if m and head < a < first and a < int(m.group(1), 16) <= first:
            p_at, p_to = a, int(m.group(1), 16)
            break
    gated_only = [t for a, t in ins if p_at is not None and p_at < a < p_to]
    full = count(common + bodies[1])
    quiet = None
    if bodies[0]:
        quiet = count(common + bodies[0])
        for k, v in count(gated_only).items():
            quiet[k] -= v
    alu = full.get("LOP3", 0) + full.get("SHF", 0)
    out = {"kernel": "v2::match_table<3,false,4>", "loop_head": hex(head), "skip_branch": hex(skip_at), "skip_target": hex(skip_to),
           "back_branch": hex(back), "fast_path_instructions": sum(full.values()), "alu_instr_per_warp_iteration": alu,
           "lop3": full.get("LOP3", 0), "shf": full.get("SHF", 0), "histogram": dict(sorted(full.items(), key=lambda kv: -kv[1])),
           "scalar_path_instructions": sum(1 for a, t in ins if skip_at < a < skip_to),
           "cc_per_warp_iteration": 127 * 32 * 4}
    if quiet:
        out["quiet_body"] = {"fast_path_instructions": sum(quiet.values()),
                             "alu_instr_per_warp_iteration": quiet.get("LOP3", 0) + quiet.get("SHF", 0),
                             "note": "the loop body a warp runs while hardly any of its threads meets a candidate (no need masks, hashed planes only)",
                             "gated_only_plane_instructions": len(gated_only)}
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    r = hot_loop(args[0] if args else None)
    print(json.dumps(r, indent=1) if "--json" in sys.argv else
          "hot loop %s..%s: %d instructions on the fast path (%d LOP3 + %d SHF = %d ALU-pipe)%s, scalar path %d instructions, %d CC per warp iteration"
          % (r["loop_head"], r["back_branch"], r["fast_path_instructions"], r["lop3"], r["shf"], r["alu_instr_per_warp_iteration"],
             "; quiet body %d (%d ALU-pipe)" % (r["quiet_body"]["fast_path_instructions"], r["quiet_body"]["alu_instr_per_warp_iteration"]) if "quiet_body" in r else "",
             r["scalar_path_instructions"], r["cc_per_warp_iteration"]))
