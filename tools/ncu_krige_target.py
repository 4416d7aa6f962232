import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, gstools_core as gc
rng = np.random.default_rng(0)
c, m = 500, 100_000
mat = rng.normal(size=(c, c)); vecs = rng.normal(size=(c, m)); cond = rng.normal(size=c)
for _ in range(2): gc.calc_field_krige_and_variance(mat, vecs, cond)
print("done", gc.dfma_peak(0, 5.0))
