"""The bench line itself is under test: a formatting slip inside one side workload once dropped cfg 3 / cfg 4 / cfg 5
from the line without failing anything (every workload is wrapped so that one failure cannot lose the headline)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _errors(obj, path=""):
    if isinstance(obj, dict):
        for k, v in obj.items():
            if k == "error":
                yield path, v
            else:
                yield from _errors(v, path + "/" + k)


@pytest.mark.gpu
def test_quick_bench_line_has_every_workload_and_no_error_entries():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--steps", "2", "--warmup", "3",
                          "--passes", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert not list(_errors(line)), list(_errors(line))
    for key in ("metric", "value", "unit", "n_gpus", "ms_per_step", "e2e", "roofline", "clocks", "gpu_launches", "config"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    for key in ("cfg3", "cfg4_n1", "cfg5_mel", "latency_b1", "latency_b4"):
        assert key in line["workloads"], (key, sorted(line["workloads"]))
    assert line["workloads"]["cfg5_mel"]["roofline"]["bound"] == "hbm"
    assert line["roofline"]["bound"] == "tensor" and 0 < line["roofline"]["frac"] < 1
