"""Diagnostic: compare tcgen05 and SIMT intermediates buffer by buffer (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neural_tensor_network_for_kb_completion_b200 import _capi
from tests.helpers import perturbed, small_problem, spread

shape = dict(Ne=500, R=11, T=3000, Nw=400, d=100, k=3, batch=2000, corrupt=10)
kb, p, batch, side, theta0 = small_problem(**shape, seed=1)
theta = np.asarray(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0), np.float32).astype(np.float64)
out = {}
for path in ("simt", "tcgen05"):
    h = _capi.Handle(p.d, p.k, p.R, p.Ne, p.Nw, p.batch_size, p.lam, wv_source="theta", gemm_path=path)
    h.set_entity_words(p.fwd_off, p.fwd_idx, p.bwd_off, p.bwd_idx, p.len_bwd)
    h.set_batch(*batch)
    c, g, n = h.eval(theta, side)
    out[path] = {nm: h.debug_fetch(nm) for nm in ("T", "M", "GA", "dWpart")}
    out[path]["grad"] = g
    print(path, c, n)
    h.close()
Kp, Np = 104, 304
for nm in ("T", "M", "GA", "dWpart", "grad"):
    a, b = out["simt"][nm], out["tcgen05"][nm]
    d = np.abs(a - b)
    i = int(d.argmax())
    print(f"{nm:7s} max|a|={np.abs(a).max():.4g} maxdiff={d.max():.4g} at {i} a={a[i]:.6g} b={b[i]:.6g} nz_a={np.count_nonzero(a)} nz_b={np.count_nonzero(b)} nan_b={np.isnan(b).sum()}")
a, b = out["simt"]["GA"].reshape(-1, Kp), out["tcgen05"]["GA"].reshape(-1, Kp)
d = np.abs(a - b)
print("GA: worst rows", np.argsort(d.max(1))[-5:], "worst cols", np.argsort(d.max(0))[-8:])
print("GA row0 simt", a[0, :8], "tc", b[0, :8])
print("GA col err profile", np.round(d.max(0)[:104], 5)[::8])
a, b = out["simt"]["dWpart"].reshape(-1, Np, Kp), out["tcgen05"]["dWpart"].reshape(-1, Np, Kp)
d = np.abs(a - b)
print("dW: chunks err", np.round(d.reshape(d.shape[0], -1).max(1), 5))
print("dW chunk0 n-profile", np.round(d[0].max(1)[::16], 5))
print("dW chunk0 p-profile", np.round(d[0].max(0)[::8], 5))
print("dW chunk0 [0,:8] simt", a[0, 0, :8], "tc", b[0, 0, :8])
