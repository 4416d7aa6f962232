"""Small workload touching every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, gstools_core as gc
rng = np.random.default_rng(0)
def modes(d, n): return rng.normal(size=(d, n)), rng.normal(size=n), rng.normal(size=n)
for d in (1, 2, 3, 5):
    k, z1, z2 = modes(d, 300)
    pos = rng.uniform(0, 9, size=(d, 1500))
    for P, L in ((0, 0), (3, 1), (1, 4), (2, 32)):
        if d > 3 and (P, L) not in ((0, 0), (1, 4)): continue
        gc.set_variant(P, L); gc.summate(k, z1, z2, pos)
        if d in (2, 3): gc.summate_incompr(k, z1, z2, pos)
gc.set_variant(0, 0)
k, z1, z2 = modes(3, 700)
gc.set_chunk_points(1024); gc.summate(k, z1, z2, rng.uniform(0, 9, size=(3, 5000))); gc.set_chunk_points(0)
gc.summate(k, z1, z2, rng.uniform(0, 9, size=(3, 400000)))          # hybrid tail tiles (P = 3)
axes3 = [np.linspace(0, 1, 9), np.linspace(0, 2, 21), np.linspace(0, 3, 37)]
axes2 = [np.linspace(0, 1, 45), np.linspace(0, 2, 133)]
gc.summate_grid(k, z1, z2, axes3); gc.summate_incompr_grid(k, z1, z2, axes3)
k2, a, b = modes(2, 100)
gc.summate_grid(k2, a, b, axes2); gc.summate_fourier_grid(a, k2, a, b, axes2)
mat = rng.normal(size=(70, 70)); vecs = rng.normal(size=(70, 333)); cond = rng.normal(size=70)
gc.calc_field_krige_and_variance(mat, vecs, cond); gc.calc_field_krige(mat, vecs, cond)
print("sanitize target done")
