"""Developer diagnostic: where a single-instance solve's wall time goes (bgg_upload_inputs / bgg_solve_resident / bgg_download_results),
against the device time of its four kernels."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, common
from common import wl
cfg_name = "a1_configuration"
cfg = wl.CONFIGS[cfg_name]
st = np.asarray(cfg["srb_init"], float)[None]
t0, ee = np.zeros(1), wl.EE_NOMINAL[None].copy()
m = common.make_gpu(cfg_name, 1, st)
z = np.zeros((1, 12 * 21 + 160))
for _ in range(20):
    m.GetRealTimeUpdate(st, t0, ee, z_out=z)
m.set_profiling(True)
T = []
for _ in range(300):
    a = time.perf_counter(); m.upload(st, t0, ee); b = time.perf_counter(); m.solve_resident(); c = time.perf_counter(); m.download(z_out=z); d = time.perf_counter()
    T.append((b - a, c - b, d - c, d - a, sum(m.last_kernel_ms().values())))
T = np.array(T) * np.array([1e3, 1e3, 1e3, 1e3, 1.0])
print("median ms: upload %.3f  solve_resident (returns after the set-up kernel) %.3f  download (waits for the solve) %.3f  total %.3f  | four kernels on the device %.3f" % tuple(np.median(T, axis=0)))
print("last solve, per kernel (ms):", {k: round(v, 3) for k, v in m.last_kernel_ms().items()}, "iterations", int(m.download()["iters"][0]))
