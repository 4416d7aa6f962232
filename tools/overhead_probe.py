import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gstools-core_b200")]
import numpy as np, gstools_core as gc
rng = np.random.default_rng(0)
k = rng.normal(size=(2, 100)); z1 = rng.normal(size=100); z2 = rng.normal(size=100)
pos = rng.uniform(0, 10, size=(2, 10000))
L = gc._load()
def med(f, n=200):
    for _ in range(20): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[n // 2] * 1e6
print("gc.summate (python wrapper + C)      : %.1f us" % med(lambda: gc.summate(k, z1, z2, pos)))
out = np.empty(10000)
args = (2, 100, 10000, k.ctypes.data, 100, 1, z1.ctypes.data, 1, z2.ctypes.data, 1, pos.ctypes.data, 10000, 1, out.ctypes.data, 0)
print("L.gsf_summate (prebuilt args, C only): %.1f us" % med(lambda: L.gsf_summate(*args)))
gc.set_profiling(True); gc.summate(k, z1, z2, pos); print("kernel_ms", gc.last_stats()["kernel_ms"]); gc.set_profiling(False)
import torch
pin = torch.from_numpy(pos).pin_memory().numpy(); pout = torch.empty(10000, dtype=torch.float64).pin_memory().numpy()
args2 = (2, 100, 10000, k.ctypes.data, 100, 1, z1.ctypes.data, 1, z2.ctypes.data, 1, pin.ctypes.data, 10000, 1, pout.ctypes.data, 0)
print("L.gsf_summate pinned in/out (zero-copy): %.1f us" % med(lambda: L.gsf_summate(*args2)))
dpos = torch.from_numpy(pos).cuda(); dout = torch.empty(10000, dtype=torch.float64, device="cuda")
args3 = (2, 100, 10000, k.ctypes.data, 100, 1, z1.ctypes.data, 1, z2.ctypes.data, 1, dpos.data_ptr(), 10000, 1, dout.data_ptr(), 0)
print("L.gsf_summate device in/out (sync)     : %.1f us" % med(lambda: L.gsf_summate(*args3)))
