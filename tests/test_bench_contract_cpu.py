"""bench.py contract, CPU side: the reference arm runs without a GPU and prints ONE JSON line with the keys the driver
reads (impl / metric / unit / value / steps / warmup / e2e / cpu_baseline / config)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*extra):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n-docs", "6000", "--dim", "64",
           "--steps", "2", "--warmup", "1", *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    # stdout carries the result line and NOTHING else (library banners such as NCCL's are routed to stderr)
    lines = out.stdout.splitlines()
    assert len(lines) == 1 and lines[0].startswith("{"), out.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = _run()
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("retrieval queries/sec") and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_hybrid_workload():
    d = _run("--workload", "hybrid")
    assert "hybrid dense+BM25 rrf" in d["config"]["workload"] and "BM25" in d["cpu_baseline"]["sample"]
