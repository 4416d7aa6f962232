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
# one-launch small path (gsf_small_kernel, both parameter-block sizes, several lane counts), then the
# repeat-mode promotion to cached records
for d, n, m in ((1, 60, 300), (2, 100, 10000), (3, 200, 2500), (3, 256, 64), (2, 33, 700)):
    k, z1, z2 = modes(d, n)
    pos = rng.uniform(0, 9, size=(d, m))
    gc.summate(k, z1, z2, pos); gc.summate_fourier(np.abs(z1), k, z1, z2, pos)
    if d in (2, 3): gc.summate_incompr(k, z1, z2, pos)
    gc.summate(k, z1, z2, pos); gc.summate(k, z1, z2, pos)
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
# degree-5 kernels (forced: the automatic rule picks them only for large problems) and the staged pageable pipeline
gc.set_poly_degree(5)
gc.summate(k, z1, z2, rng.uniform(0, 9, size=(3, 70000))); gc.summate_incompr(k, z1, z2, rng.uniform(0, 9, size=(3, 70000)))
gc.set_chunk_points(8192); gc.summate(k, z1, z2, rng.uniform(0, 9, size=(3, 70000))); gc.set_chunk_points(0)
gc.set_poly_degree(0)
# detected grid (speculative start + exact verification) and a near-grid that fails verification
g = np.stack([x.ravel() for x in np.meshgrid(*axes3, indexing="ij")]); kk, a3, b3 = modes(3, 4000)
gc.set_grid_detection(True); gc.summate(kk, a3, b3, g); g[1, 777] += 1e-9; gc.summate(kk, a3, b3, g); gc.set_grid_detection(None)
# variogram estimators (src/variogram.rs): structured / masked, unstructured (Euclid, Haversine), directional
f2 = rng.normal(size=(60, 45)); mask = rng.uniform(size=(60, 45)) < 0.2
for est in ("m", "c"):
    gc.variogram_structured(f2, est); gc.variogram_ma_structured(f2, mask, est)
for d in (1, 2, 3, 5):
    pts = rng.uniform(0, 10, size=(d, 700)); fld = rng.normal(size=(2, 700)); edges = np.linspace(0, 6, 13)
    for est in ("m", "c"):
        gc.variogram_unstructured(fld, edges, pts, est, "e")
    gc.variogram_unstructured(fld, np.array([0.0, 2.0, 1.0, 3.0, 2.5]), pts, "m", "e")     # non-monotone edges: generic kernel
    if d in (2, 3):
        dirs = rng.normal(size=(3, d)); dirs /= np.linalg.norm(dirs, axis=1)[:, None]
        gc.variogram_directional(fld, edges, pts, dirs, np.pi / 6, 1.5, False, "m")
        gc.variogram_directional(fld, edges, pts, dirs, np.pi / 6, -1.0, True, "c")
ll = np.stack([rng.uniform(-80, 80, size=500), rng.uniform(-170, 170, size=500)])
gc.variogram_unstructured(rng.normal(size=(1, 500)), np.linspace(0, 2, 9), ll, "m", "h")
print("sanitize target done")
