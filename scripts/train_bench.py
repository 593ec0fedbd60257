"""Time the Run_NTN training loop (Run_NTN.java:113-137) at Wordnet shape: per outer iteration one new batch job
(20,000 x 10) and L-BFGS with maxIters = 5 on the device.  Prints the per-iteration breakdown."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neural_tensor_network_for_kb_completion_b200 import synth
from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory
from neural_tensor_network_for_kb_completion_b200.lbfgs import LBFGSMinimizer, Opts
from neural_tensor_network_for_kb_completion_b200.ntn import NTN

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
kb = synth.make_shaped("wordnet", d=100, Nw=67447, T=20000, seed=0)
tbj = DataFactory.from_kb(kb, 20000, 10, seed=42)
t = NTN(100, kb.Ne, kb.R, kb.Nw, 20000, 3, "tanh", tbj, 1e-4, wv_source="frozen")
theta = np.asarray(t.getTheta_inital(), np.float64)
rows = []
for i in range(iters):
    t0 = time.perf_counter(); tbj.generateNewTrainingBatchJob()
    t1 = time.perf_counter(); t._sync_batch()
    t2 = time.perf_counter(); res = LBFGSMinimizer().minimize(t, theta, Opts(maxIters=5)); theta = res.minArg
    t3 = time.perf_counter()
    rows.append((t1 - t0, t2 - t1, t3 - t2, res.evals, res.firstObjVal, res.minObjVal))
    print(f"iter {i}: batch job {1e3*(t1-t0):.1f} ms, set_batch {1e3*(t2-t1):.1f} ms, L-BFGS(5) {1e3*(t3-t2):.1f} ms "
          f"({res.evals} evals), cost {res.firstObjVal:.5f} -> {res.minObjVal:.5f}", flush=True)
r = np.array([x[:4] for x in rows[2:]])
print(f"MEAN over {len(r)} iterations: batch job {1e3*r[:,0].mean():.1f} ms, set_batch {1e3*r[:,1].mean():.1f} ms, "
      f"L-BFGS {1e3*r[:,2].mean():.1f} ms ({r[:,3].mean():.1f} evals) -> {1e3*r[:,:3].sum(1).mean():.1f} ms per outer iteration; "
      f"500 iterations = {0.5*r[:,:3].sum(1).mean()*1e3:.1f} s")
