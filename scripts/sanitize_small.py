"""Small evaluation for compute-sanitizer (memcheck): both contraction paths, hub entities/words, ragged batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neural_tensor_network_for_kb_completion_b200 import _capi
from tests.helpers import perturbed, small_problem, spread

kb, p, batch, side, theta0 = small_problem(Ne=300, R=4, T=500, Nw=40, d=20, k=3, batch=400, corrupt=5, seed=1, oov_frac=0.4)
e1, rel, e2, e3 = (x.copy() for x in batch)
e3[::2] = 5
theta = np.asarray(spread(perturbed(theta0, 0.05), p), np.float32).astype(np.float64)
for path in ("simt", "tcgen05"):
    for gv in ("0", "1"):
        os.environ["NTN_GV"] = gv
        h = _capi.Handle(p.d, p.k, p.R, p.Ne, p.Nw, p.batch_size, p.lam, wv_source="theta", gemm_path=path)
        h.set_entity_words(p.fwd_off, p.fwd_idx, p.bwd_off, p.bwd_idx, p.len_bwd)
        h.set_batch(e1, rel, e2, e3)
        c, g, n = h.eval(theta, side)
        c2, g2, n2 = h.eval(theta, 1 - side)
        print(path, gv, c, n, c2, n2, float(np.abs(g).max()))
        h.close()
print("done")
