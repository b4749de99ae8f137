"""The committed bench lines carry what the bench contract asks for (keys and units), both arms name the same workload, and the ncu counts the roofline is
computed from belong to the CUDA sources in the tree (a warning, not a failure: the counts can only be re-measured on a GPU box)."""
import glob
import json
import os
import warnings

from conftest import ROOT


def latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))  # names are r02<letter>: the last is the latest
    assert files, pattern
    return json.load(open(files[-1]))


def test_bench_line_schema():
    d = latest("r02*_bench_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "configs"):
        assert k in d, k
    assert d["metric"] == "stwo_proofs_verified_per_s" and d["unit"] == "proofs/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["dtype"] == "u32" and d["scaling"] == "weak" and d["vs_baseline"] is None and "workload" in d["config"] and d["gpu_launches"] > 0
    r = d["roofline"]
    assert all(k in r for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "whole_step_frac")) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.5 < r["frac"] <= 1.0 and 0.5 < r["whole_step_frac"] <= 1.0
    c = d["cpu_baseline"]
    assert all(k in c for k in ("value", "unit", "cores", "kind", "sample")) and c["kind"] == "port"
    e = d["e2e"]
    assert all(k in e for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]  # the host link, not the kernels, bounds the end-to-end leg
    assert set(d["configs"]) >= {"c1_stark101", "c3_distinct_negatives", "c4_micro", "c5_sharded"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_names_the_same_workload():
    d, r = latest("r02*_bench_n1.json"), latest("r02*_reference_arm.json")
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["unit"] == d["unit"] and r["config"]["workload"] == d["config"]["workload"]
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0 and r["cpu_baseline"]["value"] == r["value"]


def test_pipe_counts_belong_to_this_tree():
    import bench

    doc = json.load(open(os.path.join(ROOT, "profiles", "step_pipe_counts.json")))
    assert set(doc["modes"]) == {"ref-literal", "prover-consistent"} and "stwo_merkle_kernel" in doc["modes"]["ref-literal"]["kernels"]
    if doc["csrc_sha16"] != bench.csrc_sha16():
        warnings.warn("profiles/step_pipe_counts.json was measured on other CUDA sources than the tree's: re-run tools/ncu_r02.sh + profiles/pipe_counts.py")
