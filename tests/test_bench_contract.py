"""bench.py's output contract: ONE JSON line on stdout with the keys the driver reads, for both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "gpu_launches"}


def run_bench(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]   # exactly one line on stdout, and it is JSON
    return json.loads(lines[0])


def test_reference_arm_contract(ref):
    """--impl reference: the reference's own CPU tracer (oracle/_ref), all host threads, one bounded step."""
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "ray_surface_interactions_per_s" and d["unit"] == "interactions/s"
    assert d["value"] > 1e6 and d["steps"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    # the reference arm runs oracle/_ref alone: the product library is not even loaded
    assert d["repo_libraries_loaded"] == ["oracle/_ref/libref_oracle.so"]


@pytest.mark.gpu
def test_our_arm_contract():
    d = run_bench("--steps", "5", "--warmup", "3", "--no-cpu")
    assert BASE_KEYS | {"roofline", "clocks", "strict", "parity", "e2e_full_frame", "e2e_blocking_call", "e2e_rgba8", "brackets_ms_per_step"} <= set(d)
    assert d["metric"] == "ray_surface_interactions_per_s" and d["n_gpus"] == 1 and d["steps"] == 5 and d["dtype"] == "f32"
    assert d["config"]["interactions_per_frame"] == 95944704.0 and d["config"]["jobs_per_frame"] == 87
    assert d["value"] > 1e10 and d["ms_per_step"] < 5.0          # the north star's target: a 1080p RGB flare frame in < 5 ms
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "executed_steps", "frac_executed"} <= set(r) and 0 < r["frac"] < 2 and r["peak"] > 30
    assert 0 < r["executed_steps"] < d["config"]["interactions_per_frame"] and 1500 < r["sm_clock_mhz_during_probe"] < 2200
    e = d["e2e"]   # tile-sparse: only the dirty tiles cross PCIe, and the frame in host memory is the full-frame call's
    assert 0 < e["d2h_bytes_per_step"] < 0.2 * 1920 * 1080 * 24 and e["h2d_bytes_per_step"] > 1_000_000 and 0 < e["value"] < d["value"]
    assert d["parity"]["e2e_frame_equals_full_frame_call"] is True
    assert all(d["parity"]["e2e_pipelined_frame_%d_equals_full_frame_call" % s] is True for s in range(4))   # the frames collected from flight
    assert e["value"] > d["e2e_blocking_call"]["value"] > d["e2e_full_frame"]["value"]
    assert d["e2e_full_frame"]["d2h_bytes_per_step"] == 1920 * 1080 * 24 and d["e2e_full_frame"]["value"] < e["value"]
    assert d["strict"]["ms_per_step"] < 5.0 and len(d["brackets_ms_per_step"]) == 7
    oc = d["other_configs"]   # BASELINE configs 3 and 4, one frame each, checked against the unsharded frame
    assert oc["cfg3"]["equals_single_gpu_frame"] is True and oc["cfg4"]["equals_single_gpu_frame"] is True
    assert oc["cfg3"]["interactions_per_frame"] > 4e9 and oc["cfg4"]["interactions_per_frame"] > 9e10 and oc["cfg4"]["lights"] == 64
    assert d["gpu_launches"] == 3 * 5                             # prefix + ghost + tile-finalize kernels per frame
    assert d["clocks"]["sm_max_mhz"] and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
