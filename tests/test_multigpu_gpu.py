"""Multi-GPU evidence (needs >= 2 devices; skipped on a 1-GPU box): bench.py under torchrun with
strong sharding over ranks, and the in-process gsf_set_devices path with its parity flag.
Run by `gpurun --gpus 2 -- bash tools/r2_run3.sh`; the log lands in profiles/multigpu_r2.md."""
import json
import os
import subprocess
import sys

import pytest

import gstools_core as gc
from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workload", ["c2", "c3"])
def test_bench_two_ranks_strong_sharding(workload):
    if gc.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--workload", workload, "--steps", "5",
                        "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["value"] > 0 and d["e2e"]["value"] > 0
    ip = d["in_process"]
    assert "error" not in ip, ip
    assert ip["n_devices"] == 2 and ip["parity_bit_identical_to_one_device"] is True
    print(workload, "value %.0f e2e %.0f in_process %.0f G pm/s" % (d["value"], d["e2e"]["value"], ip["value"]))
