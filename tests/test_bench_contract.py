"""bench.py's one-line JSON contract: the reference arm (host only, runs here) and the native arm (GPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*flags, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout: %r" % lines
    return json.loads(lines[0])


def test_reference_arm_line():
    """`--impl reference`: the reference's own modules (or the oracle port) on the host cores, bounded in time whatever
    --steps / --warmup are; same metric / unit / config as the native arm."""
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-budget", "6")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] == "yolov8n_640_b64_bf16" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.gpu
def test_native_arm_line():
    d = run_bench("--steps", "5", "--warmup", "3", "--cpu-seconds", "2", "--no-secondary")
    assert BASE_KEYS <= set(d) and "impl" not in d or d.get("impl") == "native"
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["unit"] == "images/s" and d["value"] > 1e5 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "yolov8n_640_b64_bf16" and "l2" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.2
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
    assert r["traffic"] is None or r["traffic"] > 0
    assert r["serial_hook"]["ms_per_forward"] > d["ms_per_step"]          # one forward in flight is slower than four
    assert set(r["kernel_ms"]) >= {"K1_reduce_planes_C3", "K2_morph_fused_C3", "K3_tile_quantize_C3"}
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 9e7 and e["d2h_bytes_per_step"] > 9e7 and e["value"] < d["value"]
    assert d["gpu_launches"] == 9 * d["steps"]
    cb = d["cpu_baseline"]
    assert cb["cores"] == 1 and cb["value"] > 0 and cb["kind"] in ("reference", "port")
    c = d["clocks"]
    assert isinstance(c.get("reasons"), list)
    if c.get("sm_mhz"):                     # nvidia-smi answered inside the (short) sampling window
        assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"]
