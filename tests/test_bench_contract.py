"""bench.py contract checks that run without a GPU: the reference arm prints one JSON line with
the keys the driver reads, on a bounded sample."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.parametrize("bp", [3, 1, 5, 4])
def test_reference_arm_prints_the_contract_line(bp):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bp", str(bp),
           "--steps", "1", "--warmup", "1", "--cpu-n", "4", "--iters", "8"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [l for l in r.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GDOF/s" and d["dtype"] == "f64" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.decode().strip() == ""
