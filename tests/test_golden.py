"""Committed golden vectors (tests/golden/ntn_small.npz, made by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them on the CPU; the CUDA path must match them on the GPU."""
import os

import numpy as np
import pytest

from oracle import ntn_ref

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ntn_small.npz"))


def problem(tag):
    d, k, R, Ne, Nw, B = (int(x) for x in G[f"{tag}_dims"])
    p = ntn_ref.NTNProblem(d=d, k=k, R=R, Ne=Ne, Nw=Nw, lam=float(G[f"{tag}_lam"]), batch_size=B,
                           fwd_off=G[f"{tag}_fwd_off"], fwd_idx=G[f"{tag}_fwd_idx"], bwd_off=G[f"{tag}_bwd_off"],
                           bwd_idx=G[f"{tag}_bwd_idx"], len_bwd=G[f"{tag}_len_bwd"], activation=tag)
    return p, tuple(G[f"{tag}_{n}"] for n in ("e1", "rel", "e2", "e3")), G[f"{tag}_theta"]


@pytest.mark.parametrize("tag", ["tanh", "sigmoid"])
@pytest.mark.parametrize("s", ["a", "b"])
def test_oracle_reproduces_golden(tag, s):
    p, batch, theta = problem(tag)
    c, g, aux = ntn_ref.compute_at(theta, p, *batch, G[f"{tag}_{s}_side"], return_aux=True)
    assert c == pytest.approx(float(G[f"{tag}_{s}_cost"]), rel=1e-13)
    assert np.allclose(g, G[f"{tag}_{s}_grad"], rtol=1e-11, atol=1e-15)
    assert aux["n_active"] == int(G[f"{tag}_{s}_nact"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["simt", "tcgen05"])
@pytest.mark.parametrize("tag", ["tanh", "sigmoid"])
@pytest.mark.parametrize("s", ["a", "b"])
def test_cuda_path_matches_golden(tag, s, path):
    from neural_tensor_network_for_kb_completion_b200 import _capi
    p, batch, theta = problem(tag)
    h = _capi.Handle(p.d, p.k, p.R, p.Ne, p.Nw, p.batch_size, p.lam, activation=tag, wv_source="theta", gemm_path=path)
    h.set_entity_words(p.fwd_off, p.fwd_idx, p.bwd_off, p.bwd_idx, p.len_bwd)
    h.set_batch(*batch)
    cost, grad, nact = h.eval(theta, G[f"{tag}_{s}_side"])
    assert nact == int(G[f"{tag}_{s}_nact"])
    assert abs(cost - float(G[f"{tag}_{s}_cost"])) <= 1e-6 + 1e-4 * abs(float(G[f"{tag}_{s}_cost"]))
    assert np.all(np.abs(grad - G[f"{tag}_{s}_grad"]) <= 1e-6 + 1e-4 * np.abs(G[f"{tag}_{s}_grad"]))
    h.close()


def test_headline_golden_vector_matches_the_oracle():
    """tests/golden/c3_fullsize.npz (the headline workload of bench.py: Freebase shape, 20,000 x 10) against a fresh
    float64 oracle run on inputs rebuilt from bench.py's own functions: the fixture is what the committed oracle and
    generator (tests/golden/make_c3_golden.py) produce.  ~30 s of CPU."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from oracle import ntn_ref
    g = np.load(bench.GOLDEN_C3)
    kb, p, maps, batch, side, theta0 = bench.headline_problem()
    theta = bench.theta_trained_like(theta0, seed=int(g["theta_seed"])).astype(np.float64)
    assert [int(x.astype(np.int64).sum()) for x in batch] == g["batch_checksum"].tolist()
    assert np.array_equal(side, g["side"])
    cost, grad, aux = ntn_ref.compute_at(theta, p, *batch, side, return_aux=True)
    assert aux["min_abs_margin"] > 2e-5                      # the guard the GPU test relies on
    assert aux["n_active"] == int(g["n_active"])
    assert cost == pytest.approx(float(g["cost"]), rel=1e-12)
    assert np.allclose(grad[g["idx"].astype(np.int64)], g["grad_at_idx"], rtol=1e-10, atol=1e-16)
    assert grad.sum() == pytest.approx(float(g["grad_sum"]), rel=1e-9)
    # and bench.py's checker accepts the oracle's own result and rejects a perturbed one
    ok = bench.check_against_golden(cost, grad, aux["n_active"], theta)
    assert ok["status"] == "pass", ok
    bad = grad.copy(); bad[int(g["idx"][1234])] += 1e-4
    assert bench.check_against_golden(cost, bad, aux["n_active"], theta)["status"] == "FAIL"
