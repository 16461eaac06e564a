"""A compiled C program linked with -lsqz_b200 (tests/c/gpu_consumer.c): the reference codec with
its search loop (squeeze.h:338-358) and token dispatch (squeeze.h:377-394) replaced by
sqz_gpu_stream_open/next/close exactly as INTEGRATION.md section 1 shows, everything else the
reference's own code included by path.  Its output must equal squeeze.compress byte for byte.
oracle/Makefile builds it into oracle/_ref/ where /root/reference is mounted."""
import os
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "oracle", "_ref", "gpu_consumer")


def test_consumer_source_follows_the_integration_guide():
    src = open(os.path.join(ROOT, "tests", "c", "gpu_consumer.c")).read()
    guide = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for call in ("sqz_gpu_stream_open", "sqz_gpu_stream_next", "sqz_gpu_stream_close", "squeeze_encode_literal",
                 "squeeze_encode_len", "squeeze_encode_pos", "squeeze_flush"):
        assert call in src and call in guide
    assert '#include "squeeze.h"' in src and "-lsqz_b200" in open(os.path.join(ROOT, "oracle", "Makefile")).read()


@pytest.mark.gpu
def test_linked_c_consumer_reproduces_the_reference_bytes(inputs, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/gpu_consumer not built (needs /root/reference at build time)")
    paths = []
    for name in ("laozi.txt", "confucius.txt", "csrc.cat"):
        p = tmp_path / name
        p.write_bytes(inputs[name].tobytes())
        paths.append(str(p))
    r = subprocess.run([BIN, *paths], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("identical") == 3, r.stdout
