"""bench.py contract, CPU side: the reference arm (the oracle port timed on host cores) prints ONE JSON line with the
keys the driver reads, honours --steps/--warmup, and needs no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0"], cwd=ROOT, text=True, stderr=subprocess.DEVNULL, timeout=600)
    lines = [l for l in out.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "audio_seconds_per_second" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["scaling"] == "weak"
    assert d["value"] > 0 and abs(d["value"] - 8.0 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]   # 2 x 4 s per step
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None
