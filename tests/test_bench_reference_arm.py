"""`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) works without a GPU and prints one JSON
line with the contract's keys, on the same metric / config as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    import bench
    assert d["impl"] == "reference" and d["unit"] == "SNPs/s" and d["higher_is_better"] is True
    assert d["metric"] == bench.WORKLOADS["c3"]["metric"]
    assert d["config"] == bench.config_of("c3", bench.WORKLOADS["c3"]["snps"])
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": "SNPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
