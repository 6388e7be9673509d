"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU restatement timed on the host
cores) prints exactly ONE JSON line on stdout with the keys the driver reads, and names the same workload as our arm."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold one JSON line, got {len(lines)}"
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same metric / workload strings as our arm (the driver compares the two lines)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('metric="rays/sec (render, 2x128 samples)"') >= 2
    assert len(re.findall(r"workload=WORKLOAD_C2", src)) == 2
    assert d["config"]["workload"].startswith("C2 full-frame render 1920x1280")
