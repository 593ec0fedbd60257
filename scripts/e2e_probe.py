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
print(cfg, "per-call ms:", [round(x, 2) for x in ts], "nproc", os.cpu_count(), "load", os.getloadavg())
