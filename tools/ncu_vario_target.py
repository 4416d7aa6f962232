"""ncu target: one isotropic, one directional and one structured variogram call (see profiles/)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, gstools_core as gc
rng = np.random.default_rng(1)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "iso2"):
    m = 200000
    pos = rng.uniform(0.0, 1000.0, (2, m)); f = rng.normal(size=(1, m))
    gc.variogram_unstructured(f, np.linspace(0.0, 300.0, 31), pos)
if which in ("all", "iso3"):
    m = 100000
    pos = rng.uniform(0.0, 1000.0, (3, m)); f = rng.normal(size=(1, m))
    gc.variogram_unstructured(f, np.linspace(0.0, 300.0, 31), pos)
if which in ("all", "dir3"):
    m = 50000
    pos = rng.uniform(0.0, 1000.0, (3, m)); f = rng.normal(size=(1, m))
    gc.variogram_directional(f, np.linspace(0.0, 300.0, 21), pos, np.eye(3), np.pi / 8, 50.0)
if which in ("all", "struct"):
    gc.variogram_structured(rng.normal(size=(4000, 4000)))
print("done")
