"""Dump the cycle-stamp timeline of CTA (0,0,0) of each tcgen05 contraction (GPU box; NTN_TC_TIMELINE=1)."""
import os, sys
os.environ["NTN_TC_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sys.argv = ["bench.py"]
import bench
from neural_tensor_network_for_kb_completion_b200 import _capi
from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory

kb = bench.build_problem(1)
tbj = DataFactory.from_kb(kb, bench.BATCH, bench.CORRUPT, seed=42)
h = _capi.Handle(bench.D, bench.K_SLICES, kb.R, kb.Ne, kb.Nw, bench.BATCH, bench.LAMBDA, wv_source="theta", gemm_path="tcgen05")
h.set_entity_words(*tbj.entityWordMaps())
h.set_batch(*tbj.generateNewTrainingBatchJob())
rng = np.random.default_rng(3)
theta = np.concatenate([rng.random(kb.R * 3 * 100 * 100) * 0.07, np.full(kb.R * 600, 0.4), np.full(kb.R * 3, 0.01),
                        np.full(kb.R * 3, 0.8), kb.wv.astype(np.float64).ravel()])
side = np.ones(kb.R, np.uint8)
for _ in range(5):
    h.eval(theta, side)
tl = h.debug_fetch("tc_timeline").reshape(3, 2, 128)
for m, name in enumerate(("fwd", "dw", "ge")):
    p = tl[m, 0]; q = tl[m, 1]
    p = p[p >= 0]; q = q[q >= 0]
    print(name, "producer stamps (cycles):", np.round(p).astype(int).tolist())
    print(name, "mma issuer stamps      :", np.round(q).astype(int).tolist())
