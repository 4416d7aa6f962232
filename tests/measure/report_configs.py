"""Run bench.py on each of the five BASELINE.json configs (1 GPU) and write profiles/configs_r2.md:
kernel-only and end-to-end G point*modes/s, roofline fraction, CPU baseline, structured-grid path,
and max|delta|/sigma against the oracle on a point subset."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np

STEPS = {"c1": 200, "c2": 200, "c3": 200, "c4": 20, "c5": 5}
rows = []
for cfg in ("c1", "c2", "c3", "c4", "c5"):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", cfg, "--steps", str(STEPS[cfg]),
                          "--warmup", "3", "--cpu-seconds", "8"], capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    rows.append((cfg, d))
    print(cfg, "value", d["value"], "e2e", d["e2e"]["value"], flush=True)

import gstools_core as gc, oracle
from gstools_core import workloads
gc.set_grid_detection(False)
errs = {}
for cfg in ("c1", "c2", "c3", "c4", "c5"):
    w = workloads.make(cfg)
    got = getattr(gc, w["kind"])(*w["args"])
    m = w["m"]
    idx = np.unique(np.concatenate([np.arange(0, m, max(1, m // 4096)), np.arange(min(m, 512)), np.arange(max(0, m - 512), m)]))
    ref = getattr(oracle, w["kind"])(*workloads.subset_points(w, idx)["args"], oracle.max_threads())
    sub = got[:, idx] if got.ndim == 2 else got[idx]
    errs[cfg] = float(np.max(np.abs(sub - ref)) / np.std(ref))
    del got

with open(os.path.join(ROOT, "profiles", "configs_r2.jsonl"), "w") as f:
    f.write("\n".join(json.dumps(d) for _, d in rows) + "\n")
with open(os.path.join(ROOT, "profiles", "configs_r2.md"), "w") as f:
    f.write("# The five BASELINE.json configs on one B200 (round 2)\n\n")
    f.write("`python tests/measure/report_configs.py` = `bench.py --workload cN` per config (general point x mode kernel, grid detection off\n"
            "for kernel / e2e / roofline; `default API` = what the plain call does on this input, i.e. the structured-grid path for C2-C5)\n"
            "plus a parity check of the full-size result against the CPU oracle on a strided point subset.  G pm/s = 1e9 point*modes per\n"
            "second.  e2e = `gstools_core.<function>(pageable numpy arrays) -> host ndarray`.  Raw bench lines: `profiles/configs_r2.jsonl`.\n\n")
    f.write("| cfg | function | d | modes | points | kernel | kernel G pm/s | FP64 roofline frac (of measured DFMA peak) | e2e pageable: ms, G pm/s | e2e caller-pinned ms | default API ms (grid path) | CPU oracle G pm/s (cores) | e2e / CPU | max abs diff / sigma vs oracle |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for cfg, d in rows:
        c = d["config"]
        g = d.get("structured_grid")
        cb = d.get("cpu_baseline", {})
        f.write("| %s | %s | %d | %d | %d | %s | %.0f | %.3f | %.4g, %.0f | %.4g | %s | %.3f (%s) | %.0fx | %.2e |\n" % (
            cfg.upper(), c["kind"], c["dim"], c["modes"], c["points_total"], d["roofline"]["kernel"].replace("gsf_sum_kernel", ""), d["value"],
            d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["pinned_input"]["ms_per_step"],
            ("%.4g" % g["e2e"]["ms_per_step"]) if g else "= e2e (points on a line: no grid)",
            cb.get("value", float("nan")), cb.get("cores", "?"), d["e2e"]["value"] / cb.get("value", float("nan")), errs[cfg]))
    f.write("\nCPU oracle: OpenMP C restatement of the Rayon path (`oracle/`), all host cores, bounded sample scaled linearly in points\n"
            "(C4/C5: extrapolated).  C1 is launch-latency bound (1e6 point*modes = 1 us of FP64 work): see profiles/latency_c1_r2.md.\n")
print(open(os.path.join(ROOT, "profiles", "configs_r2.md")).read()[:3000])
