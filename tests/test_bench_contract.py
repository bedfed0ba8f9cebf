"""bench.py's reference arm (`--impl reference`): runs on host cores only, so its record contract can be checked here -- one JSON
line on stdout with the keys the driver reads, the unmodified reference loaded from baseline/_ref where that exists (the C port,
kind "port", otherwise)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "segment-timesteps/sec PERLND SNOW+PWATER fp64"
    assert d["unit"] == "segment-timesteps/s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["n_gpus"] == 1
    assert d["value"] > 1e5 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("test10 PERLND SNOW+PWATER replicated to 100000 segments")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "hsp2", "hsp2")):
        assert cb["kind"] == "reference" and ("baseline/_ref" in cb["loaded_from"] or "/root/reference" in cb["loaded_from"])
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
