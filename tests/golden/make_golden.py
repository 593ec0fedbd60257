#!/usr/bin/env python
"""Generates tests/golden/ntn_small.npz: a small seeded problem with the float64 oracle's cost, gradient and
n_active for both corrupt-side assignments (tanh) and one sigmoid case.

The reference is Java and cannot be imported or run here (no JVM), so these vectors are produced by the oracle
restatement (oracle/ntn_ref.py) -- they pin the oracle and the CUDA path against silent drift, not against
ND4j (parity unpinned).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ntn_ref                                          # noqa: E402
from tests.helpers import perturbed, small_problem, spread          # noqa: E402


def main():
    out = {}
    for tag, act in (("tanh", "tanh"), ("sigmoid", "sigmoid")):
        kb, p, batch, side, theta0 = small_problem(Ne=60, R=3, T=120, Nw=40, d=10, k=2, batch=80, corrupt=3, seed=21,
                                                   activation=act, lam=1e-3)
        theta = np.asarray(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0), np.float32).astype(np.float64)
        for sname, sd in (("a", side), ("b", 1 - side)):
            c, g, aux = ntn_ref.compute_at(theta, p, *batch, sd, return_aux=True)
            assert aux["min_abs_margin"] > 1e-4
            out[f"{tag}_{sname}_side"] = np.asarray(sd, np.uint8)
            out[f"{tag}_{sname}_cost"] = np.float64(c)
            out[f"{tag}_{sname}_grad"] = g
            out[f"{tag}_{sname}_nact"] = np.int64(aux["n_active"])
        out[f"{tag}_theta"] = theta
        for nm, arr in zip(("e1", "rel", "e2", "e3"), batch):
            out[f"{tag}_{nm}"] = arr
        for nm in ("fwd_off", "fwd_idx", "bwd_off", "bwd_idx", "len_bwd"):
            out[f"{tag}_{nm}"] = getattr(p, nm)
        out[f"{tag}_dims"] = np.array([p.d, p.k, p.R, p.Ne, p.Nw, p.batch_size], np.int64)
        out[f"{tag}_lam"] = np.float64(p.lam)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ntn_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
