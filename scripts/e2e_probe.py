"""Probe the host-buffer path (ntn_eval): per-call wall time over repeated calls, with its parts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sys.argv = ["bench.py"]
import bench
from neural_tensor_network_for_kb_completion_b200 import _capi
from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory
cfg = os.environ.get("CFG", "c3")
bench.SHAPE, bench.BATCH, bench.CONFIG_NAME = bench.CONFIGS[cfg]
kb = bench.build_problem(1)
tbj = DataFactory.from_kb(kb, bench.BATCH, bench.CORRUPT, seed=42)
h = _capi.Handle(bench.D, bench.K_SLICES, kb.R, kb.Ne, kb.Nw, bench.BATCH, bench.LAMBDA, wv_source="theta")
h.set_entity_words(*tbj.entityWordMaps())
h.set_batch(*tbj.generateNewTrainingBatchJob())
theta = np.random.default_rng(0).standard_normal(h.P) * 0.05
side = np.ones(kb.R, np.uint8)
g = np.empty(h.P)
ts = []
for i in range(25):
    t0 = time.perf_counter(); h.eval(theta, side, grad_out=g); ts.append(1e3 * (time.perf_counter() - t0))
knobs = {k: os.environ.get(k) for k in ("NTN_NT_STORES", "NTN_CONVERT_THREADS", "NTN_XFER_CHUNK_LOG2") if k in os.environ}
print(cfg, knobs, "median ms %.3f min %.3f" % (np.median(ts[3:]), min(ts)), "nproc", os.cpu_count(), "load", os.getloadavg(), flush=True)
if os.environ.get("PROBE_PARTS"):
    import torch
    lib = h.lib
    d = torch.empty(h.Pint, dtype=torch.float32, device="cuda")
    for name, fn in (("upload_theta (convert + H2D + layout)", lambda: h.upload_theta(theta, d.data_ptr())),
                     ("download_theta (layout + D2H + convert)", lambda: h.download_theta(d.data_ptr(), out=g))):
        tt = []
        for i in range(15):
            t0 = time.perf_counter(); fn(); tt.append(1e3 * (time.perf_counter() - t0))
        print("  ", name, "median ms %.3f" % np.median(tt[3:]))
    f32 = np.empty(h.P, np.float32)
    for name, fn in (("convert f64->f32 only", lambda: lib.ntn_host_convert_f64_f32(theta.ctypes.data, f32.ctypes.data, h.P)),
                     ("convert f32->f64 only", lambda: lib.ntn_host_convert_f32_f64(f32.ctypes.data, g.ctypes.data, h.P))):
        tt = []
        for i in range(15):
            t0 = time.perf_counter(); fn(); tt.append(1e3 * (time.perf_counter() - t0))
        print("  ", name, "median ms %.3f" % np.median(tt[3:]))
    pin = torch.empty(h.P, dtype=torch.float32).pin_memory()
    dv = torch.empty(h.P, dtype=torch.float32, device="cuda")
    for name, fn in (("pinned H2D %.1f MB" % (h.P * 4 / 1e6), lambda: dv.copy_(pin, non_blocking=True)),
                     ("pinned D2H", lambda: pin.copy_(dv, non_blocking=True))):
        tt = []
        for i in range(10):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); tt.append(1e3 * (time.perf_counter() - t0))
        print("  ", name, "median ms %.3f" % np.median(tt[2:]))
