"""Times the variogram estimators on the GPU next to the CPU oracle (bounded oracle sizes: the
reference algorithm is O(bins * M^2)).  Usage: python tests/measure/variogram_probe.py [out.md]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gstools-core_b200"))
import gstools_core as gc  # noqa: E402
import oracle  # noqa: E402


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t)
    return min(ts), r


def main():
    rows = []
    rng = np.random.default_rng(1)
    thr = oracle.max_threads()
    gc.set_profiling(True)
    for d, m, nb, m_cpu in [(2, 20000, 30, 4000), (3, 100000, 30, 4000), (2, 400000, 30, 0)]:
        pos = rng.uniform(0.0, 1000.0, (d, m))
        f = rng.normal(size=(1, m))
        edges = np.linspace(0.0, 300.0, nb + 1)
        gc.variogram_unstructured(f, edges, pos)
        t_gpu, (g, c) = best(lambda: gc.variogram_unstructured(f, edges, pos))
        kms = gc.last_stats()["kernel_ms"]
        pairs = m * (m - 1) / 2
        line = "| unstructured d=%d M=%d bins=%d | %.2f ms (kernel %.2f ms) | %.1f G pairs/s |" % (
            d, m, nb, t_gpu * 1e3, kms, pairs / (kms * 1e-3) / 1e9)
        if m_cpu:
            sub = slice(0, m_cpu)
            t_cpu, (go, co) = best(lambda: oracle.variogram_unstructured(f[:, sub], edges, pos[:, sub], "m", "e", thr), 1)
            gs, cs = gc.variogram_unstructured(f[:, sub], edges, pos[:, sub])
            assert np.array_equal(cs, co) and np.allclose(gs, go, rtol=1e-12, atol=0)
            cpu_pairs = m_cpu * (m_cpu - 1) / 2
            line += " oracle %d threads on M=%d: %.2f s = %.4f G pairs/s |" % (thr, m_cpu, t_cpu, cpu_pairs / t_cpu / 1e9)
        else:
            line += " - |"
        rows.append(line)
        print(line, flush=True)
    # directional
    for d, m, nb in [(3, 50000, 20), (2, 100000, 20)]:
        pos = rng.uniform(0.0, 1000.0, (d, m))
        f = rng.normal(size=(1, m))
        edges = np.linspace(0.0, 300.0, nb + 1)
        direction = np.eye(d)
        gc.variogram_directional(f, edges, pos, direction, np.pi / 8, 50.0)
        t_gpu, _ = best(lambda: gc.variogram_directional(f, edges, pos, direction, np.pi / 8, 50.0))
        kms = gc.last_stats()["kernel_ms"]
        line = "| directional d=%d M=%d bins=%d dirs=%d bw=50 | %.2f ms (kernel %.2f ms) | %.1f G pairs/s | - |" % (
            d, m, nb, d, t_gpu * 1e3, kms, m * (m - 1) / 2 / (kms * 1e-3) / 1e9)
        rows.append(line)
        print(line, flush=True)
    # Haversine (degrees on the unit sphere)
    m, nb = 50000, 20
    pos = np.stack([rng.uniform(-80.0, 80.0, m), rng.uniform(-180.0, 180.0, m)])
    f = rng.normal(size=(1, m))
    edges = np.linspace(0.0, 1.0, nb + 1)
    gc.variogram_unstructured(f, edges, pos, "m", "h")
    t_gpu, _ = best(lambda: gc.variogram_unstructured(f, edges, pos, "m", "h"))
    kms = gc.last_stats()["kernel_ms"]
    line = "| unstructured Haversine M=%d bins=%d | %.2f ms (kernel %.2f ms) | %.1f G pairs/s | - |" % (
        m, nb, t_gpu * 1e3, kms, m * (m - 1) / 2 / (kms * 1e-3) / 1e9)
    rows.append(line)
    print(line, flush=True)
    # structured
    for shape, shape_cpu in [((2000, 2000), (500, 2000)), ((8000, 4000), None)]:
        fs = rng.normal(size=shape)
        gc.variogram_structured(fs)
        t_gpu, g = best(lambda: gc.variogram_structured(fs))
        kms = gc.last_stats()["kernel_ms"]
        pairs = shape[0] * (shape[0] - 1) / 2 * shape[1]
        line = "| structured %dx%d | %.2f ms (kernel %.2f ms) | %.1f G pairs/s |" % (
            shape[0], shape[1], t_gpu * 1e3, kms, pairs / (kms * 1e-3) / 1e9)
        if shape_cpu:
            fc = fs[:shape_cpu[0]]
            t_cpu, go = best(lambda: oracle.variogram_structured(fc, "m", thr), 1)
            assert np.allclose(gc.variogram_structured(fc), go, rtol=1e-12, atol=0)
            cp = shape_cpu[0] * (shape_cpu[0] - 1) / 2 * shape_cpu[1]
            line += " oracle %d threads on %dx%d: %.2f s = %.3f G pairs/s |" % (thr, shape_cpu[0], shape_cpu[1], t_cpu, cp / t_cpu / 1e9)
        else:
            line += " - |"
        rows.append(line)
        print(line, flush=True)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as fh:
            fh.write("| case | GPU call (host arrays in/out) | kernel rate | CPU oracle |\n|---|---|---|---|\n")
            fh.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
