"""The bench.py JSON line contract, checked on the lines banked under profiles/ (no GPU needed): every key the driver and the judge read
is present, typed and self-consistent.  Guards against a bench.py edit that silently drops or renames a key."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles", "r1")


def _line(name):
    with open(os.path.join(P, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n_gpus", [("bench_cfg3_r1_final_default.json", 1), ("bench_cfg3_r1_final_dp1.json", 1), ("bench_cfg3_r1_final_dp2.json", 2),
                                         ("../r2/bench_cfg3.json", 1), ("../r2/bench_cfg3_final_check.json", 1), ("../r2/bench_cfg3_n2.json", 2),
                                         ("../r2/bench_cfg3_n4.json", 4), ("../r2/bench_cfg3_n8.json", 8)])
def test_own_arm_line(name, n_gpus):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["metric"].startswith("MIDI sequences/sec") and d["unit"] == "sequences/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == n_gpus and d["scaling"] == "weak" and d["dtype"] == "bf16" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 10
    assert "cfg3" in d["config"]["workload"] and "seq_len=256 hidden=512" in d["config"]["workload"] and "l2" in d["config"]
    # value = whole-job sequences / max-over-ranks step time
    assert abs(d["value"] - 512 * n_gpus / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.8 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["step"]["frac"] < 1 and r["step"]["train_flops_per_seq"] == 13395492864


def test_default_line_has_cpu_baseline_and_measured_traffic():
    d = _line("bench_cfg3_r1_final_default.json")
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["unit"] == "sequences/s" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["roofline"]["traffic"] is not None


def test_reference_arm_line():
    d = _line("bench_cfg3_r1_final_reference.json")
    assert d["impl"] == "reference" and d["metric"].startswith("MIDI sequences/sec") and d["unit"] == "sequences/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]


def test_reference_default_gru_line():
    """The reference's own default configuration (GRU, settings.py:108-115,155) through the cluster-resident GRU kernels."""
    d = _line("../r2/bench_refdefault_gru.json")
    assert "refdefault" in d["config"]["workload"] and "GRU" in d["config"]["workload"] and d["dtype"] == "bf16"
    assert abs(d["value"] - 256 / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    assert d["roofline"]["kernel"].startswith("gru_cluster_") and d["e2e"]["h2d_bytes_per_step"] > 0 and d["cpu_baseline"]["kind"] == "port"
    s = _line("../r2/bench_refdefault_gru_streamed.json")
    assert d["value"] > 8 * s["value"]                 # one launch per recurrence instead of four per step


def test_flop_model_matches_the_survey():
    import bench
    fwd, rec = bench.flops_per_seq(256, 512, 256, "teacher_forced")
    assert 3 * fwd == 13395492864                      # 13.3955 GFLOP per sequence per train step at cfg3
    fwd2, _ = bench.flops_per_seq(64, 256, 100, "teacher_forced")
    assert abs(3 * fwd2 / 1e9 - 0.8778) < 1e-3         # cfg2
    fwd1, _ = bench.flops_per_seq(16, 64, 16, "teacher_forced")
    assert abs(3 * fwd1 / 1e6 - 17.2) < 0.1            # cfg1
