"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the
required keys, and the synthetic workloads are what SURVEY.md section 8 d2 describes."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from gstools_core import workloads


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")          # torchrun exports this; the arm must ignore it
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--scale", "0.002"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["scaling"] == "strong"
    # the `ours` arm must print the very same config dict (the driver compares them)
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    ns = argparse.Namespace(workload="c5", scale=0.002, gpus=1)
    assert d["config"] == bench.config_for(ns, d["config"]["points_total"])


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--scale", "0.002"], capture_output=True, text=True, env=env,
                       timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_shapes_and_grids():
    sizes = {"c1": (2, 100, 10_000), "c2": (3, 1000, 1_000_000), "c3": (3, 1000, 1_000_000)}
    for cfg, (d, n, m) in sizes.items():
        w = workloads.make(cfg)
        assert (w["d"], w["n"], w["m"]) == (d, n, m) and w["args"][-1].shape == (d, m)
    for cfg in ("c2", "c3", "c4", "c5"):
        w = workloads.make(cfg, 0.001)
        g = np.meshgrid(*w["axes"], indexing="ij")
        assert np.array_equal(np.stack([x.ravel() for x in g]), w["args"][-1])   # axes expand to pos exactly
    w4 = workloads.make("c4", 0.001)
    assert w4["kind"] == "summate_fourier" and w4["n"] == 10_000 and len(w4["args"]) == 5
    w5 = workloads.make("c5", 1e-6)
    assert w5["n"] == 10_000 and w5["d"] == 3
    assert workloads.W_EXEC["c2"] == 14 and workloads.W_SURVEY["c2"] == 23
