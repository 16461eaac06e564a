#!/usr/bin/env python
"""Regenerate tests/golden/ from the mounted reference (run HERE, where
/root/reference exists; the GPU box only reads the committed outputs).

Writes
  fixtures.tar.xz  the six data files of /root/reference/test (inputs only)
  golden.json      per input and window: what the UNMODIFIED reference
                   (oracle/_ref) produced -- compressed size, FNV-1a-64 of the
                   bitstream in memory mode and in callback/file mode, token count
                   and digest (ref_tokens) -- plus digests of the full match table
                   computed by oracle/sqz_oracle.c after that restatement was
                   checked against the reference on the same input.
"""
import hashlib
import io
import json
import os
import sys
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
REF_TEST = REF + "/test"
# csrc.cat stands in for the missing test/sqlite3.c
CSRC_CAT = ["src/sqz.c", "shl/sqz/sqz.h", "test.c", "bst.c", "shl.c", "inc/rt/rt.h", "inc/rt/rt_generics.h",
            "inc/rt/rt_generics_test.h", "inc/rt/ustd.h", "inc/rt/fileio.h", "inc/sqz/sqz.h",
            "attic/map_experiment/bitstream.h", "attic/map_experiment/file.h", "attic/map_experiment/huffman.h",
            "attic/map_experiment/map.h", "attic/map_experiment/squeeze.h", "attic/map_experiment/test.c"]

from oracle import Oracle, Reference, build  # noqa: E402
from sqz_b200 import corpus  # noqa: E402


def pack():
    def reset(ti):
        ti.uid = ti.gid = 0
        ti.uname = ti.gname = ""
        ti.mtime = 0
        ti.mode = 0o644
        return ti

    with tarfile.open(os.path.join(HERE, "fixtures.tar.xz"), "w:xz", preset=9) as tf:
        for n in corpus.ORDER:
            tf.add(os.path.join(REF_TEST, n), arcname=n, filter=reset)
        cat = b"".join(open(os.path.join(REF, f), "rb").read() for f in CSRC_CAT)
        assert len(cat) == 179548, len(cat)
        ti = reset(tarfile.TarInfo("csrc.cat"))
        ti.size = len(cat)
        tf.addfile(ti, io.BytesIO(cat))


def blob_sha1(b: bytes) -> str:
    return hashlib.sha1(b"blob %d\0" % len(b) + b).hexdigest()


def main():
    build()
    pack()
    o, r = Oracle.get(), Reference.get()
    inputs = {k: np.frombuffer(v, dtype=np.uint8) for k, v in corpus.kat_inputs().items()}
    inputs.update(corpus.all_files())
    out = {"_about": "made by tests/golden/make_golden.py from oracle/_ref (unmodified reference) "
                     "and oracle/sqz_oracle.c; FNV-1a-64 digests, hex",
           "inputs": {}}
    for name, d in inputs.items():
        e = {"bytes": int(d.size), "git_blob": blob_sha1(d.tobytes())[:12], "win": {}}
        for wb in (10, 15):
            comp = r.compress(d, wb)
            comp_f = r.compress(d, wb, file_mode=True)
            assert r.decompress(comp) == d.tobytes()
            rt = r.tokens(comp)
            ot = o.tokens(d, 1 << wb)
            assert rt.size == ot.size and (rt == ot).all(), (name, wb)
            ln, ds = o.match_table(d, 1 << wb)
            tt, end = o.tokens_from_table(d, ln, ds)
            assert end == d.size and (tt == rt).all(), (name, wb)
            assert r.encode_tokens(tt, d.size, wb) == comp
            w = {
                "compressed_bytes": len(comp),
                "fnv_mem": "%016x" % o.fnv(np.frombuffer(comp, np.uint8)),
                "fnv_file": "%016x" % o.fnv(np.frombuffer(comp_f, np.uint8)),
                "tokens": int(rt.size),
                "matches": int((rt >> 16 != 0).sum()),
                "fnv_tokens": "%016x" % o.fnv(rt),
                "fnv_len": "%016x" % o.fnv(ln),
                "fnv_dist": "%016x" % o.fnv(ds),
            }
            if d.size <= 64:
                w["hex_mem"] = comp.hex()
                w["token_list"] = [int(x) for x in rt]
            e["win"][str(wb)] = w
            print(name, wb, w["compressed_bytes"], w["fnv_mem"], w["tokens"], flush=True)
        out["inputs"][name] = e
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
