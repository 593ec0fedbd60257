#!/usr/bin/env python
"""Golden vector of the HEADLINE workload (BASELINE.json configs[2], exactly what bench.py times): the float64 oracle's
cost, n_active and a strided sample of the gradient of one full 20,000 x 10 evaluation.

    python tests/golden/make_c3_golden.py          # ~1-2 min of CPU, writes tests/golden/c3_fullsize.npz

The problem is built by bench.py's own functions, so the GPU test (tests/test_gpu_parity.py::
test_headline_config_full_size_against_golden) and bench.py's `gpu_vs_oracle_f64` check evaluate the very same inputs.
theta = theta0 + N(0, 0.05^2) with the first seed (starting at bench.py's 11) for which the float64 oracle sees no
hinge margin within 2e-5 of zero: the strict FP32 comparison of NTN.java:213 may legitimately flip such a column, and a
flip moves gradient entries by O(1/batchSize) = 5e-5, far outside the 1e-6 + 1e-4|ref| tolerance -- the same guard as
tests/test_gpu_parity.py::check.  The oracle is the restatement of /root/reference/src/NTN.java:92-443 in
oracle/ntn_ref.py (parity unpinned: the reference cannot run here; tools/DumpGolden.java produces the same file layout
from the real reference on a machine with a JVM).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402
from oracle import data_ref, ntn_ref                             # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c3_fullsize.npz")
STRIDE = 37
NEAR_TIE = 2e-5


def margins_f64(theta, p, batch, side):
    """Hinge margins of every column (forward pass only, float64) -- used to screen theta seeds for near-ties."""
    e1, rel, e2, e3 = batch
    W, V, b, U, WV = ntn_ref.stack_to_parameters(theta, p)
    E = data_ref.create_entity_vectors(WV, p.fwd_off, p.fwd_idx, np.float64)
    out = np.empty(len(e1))
    for r in range(p.R):
        cols = np.nonzero(rel == r)[0]
        if cols.size == 0:
            continue
        E1, E2, E3 = E[:, e1[cols]], E[:, e2[cols]], E[:, e3[cols]]
        E1n, E2n = (E1, E3) if side[r] else (E3, E2)

        def score(A, C):
            pre = np.stack([np.sum(A * (W[r, s] @ C), axis=0) for s in range(p.k)]) + V[r].T @ np.vstack([A, C]) + b[r][:, None]
            return U[r] @ np.tanh(pre)
        out[cols] = score(E1, E2) + 1.0 - score(E1n, E2n)
    return out


def main():
    kb, p, maps, batch, side, theta0 = bench.headline_problem()
    seed, theta = None, None
    for s in range(bench.THETA_SEED_DEFAULT, bench.THETA_SEED_DEFAULT + 200):
        th = bench.theta_trained_like(theta0, seed=s).astype(np.float64)
        m = margins_f64(th, p, batch, side)
        print(f"theta seed {s}: min |margin| {np.abs(m).min():.3e}, {int((np.abs(m) < NEAR_TIE).sum())} near-ties", flush=True)
        if np.abs(m).min() > NEAR_TIE:
            seed, theta = s, th
            break
    assert seed is not None
    t0 = time.time()
    cost, grad, aux = ntn_ref.compute_at(theta, p, *batch, side, return_aux=True)
    print(f"float64 oracle: {time.time() - t0:.1f} s, cost {cost:.12f}, n_active {aux['n_active']}, "
          f"min |margin| {aux['min_abs_margin']:.3e}")
    assert aux["min_abs_margin"] > NEAR_TIE
    offs = np.array(list(p.offsets()) + [p.P], np.int64)
    idx = np.arange(0, p.P, STRIDE, dtype=np.int64)
    small = np.arange(offs[1], offs[4], dtype=np.int64)            # V, b, U in full
    idx = np.unique(np.concatenate([idx, small]))
    np.savez_compressed(
        OUT, theta_seed=seed, cost=cost, n_active=aux["n_active"], min_abs_margin=aux["min_abs_margin"],
        P=p.P, offsets=offs, idx=idx.astype(np.int32), grad_at_idx=grad[idx],
        block_l2=np.array([np.linalg.norm(grad[a:b]) for a, b in zip(offs[:-1], offs[1:])]),
        grad_sum=grad.sum(), grad_abs_sum=np.abs(grad).sum(),
        theta_checksum=float(np.dot(theta[::STRIDE], np.arange(len(theta[::STRIDE])) % 7 + 1.0)),
        batch_checksum=np.array([int(x.astype(np.int64).sum()) for x in batch], np.int64),
        side=np.asarray(side, np.uint8))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
