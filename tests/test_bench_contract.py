"""bench.py pieces that can be checked without a GPU."""
import importlib.util
import os

from conftest import ROOT


def load_bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_candidate_compare_count_is_the_survey_formula():
    b = load_bench()
    W = 32768
    # whole input N >= W: (W-1)(N - W/2); confucius.txt: 1,682,618,217 exactly
    assert b.cc_count(0, 67735, W - 1) == 32767 * (67735 - 16384) == 1682618217
    for g0, n in [(0, 10), (0, 40000), (5, 100), (32760, 20), (10**6, 12345)]:
        assert b.cc_count(g0, n, W - 1) == sum(min(i, W - 1) for i in range(g0, g0 + n))
    assert b.cc_count(1 << 30, 1 << 20, W - 1) == (1 << 20) * 32767          # away from the start


def test_workload_names_the_baseline_config():
    b = load_bench()

    class A:
        size, gpus = 1 << 30, 1
    cfg = b.workload_config(A)
    assert "synthetic corpus" in cfg["workload"] and cfg["window"] == 32768
    assert (cfg["min_len"], cfg["max_len"], cfg["max_dist"]) == (3, 257, 32767)
    assert b.METRIC == "match_search_input_MBps"
