"""CPU check of bench.py's reference arm (the only leg of bench.py that runs without a GPU): one JSON line
on stdout with the contract's keys, rank > 0 silent, both workloads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
        "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("extra", [["--grid", "12"], ["--grid", "40", "--workload", "cfg2"],
                                   ["--grid", "20", "--workload", "cfg5"]])
def test_reference_arm_prints_one_contract_line(extra):
    r = _run(extra)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "V-cycles/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 2 and "workload" in d["config"]
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--grid", "12", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
