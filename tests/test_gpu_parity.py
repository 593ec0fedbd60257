"""GPU parity: the CUDA path, called through the C-ABI (ntn_eval & friends), against the
float64 oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): cost and every gradient element within 1e-4 relative and
1e-6 absolute, i.e. |x - ref| <= 1e-6 + 1e-4*|ref|; n_active (integer) exact.  Each case first
checks that the oracle sees no near-tie in the hinge (|margin| > 2e-5), so an FP32 flip of the
strict comparison NTN.java:213 cannot hide behind the tolerance.
"""
import numpy as np
import pytest

from neural_tensor_network_for_kb_completion_b200 import _capi
from oracle import data_ref, ntn_ref
from oracle.java_random import JavaRandom
from tests.helpers import perturbed, small_problem, spread

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-6


def make_handle(p, wv_source="theta", gemm_path="auto", kb=None, reg_scale=1.0):
    h = _capi.Handle(p.d, p.k, p.R, p.Ne, p.Nw, p.batch_size, p.lam, activation=p.activation,
                     wv_source=wv_source, gemm_path=gemm_path, reg_scale=reg_scale)
    h.set_entity_words(p.fwd_off, p.fwd_idx, p.bwd_off, p.bwd_idx, p.len_bwd)
    if wv_source == "frozen":
        h.set_frozen_wv(kb.wv)
    return h


def f32(theta):
    return np.asarray(theta, np.float32).astype(np.float64)


def check(h, p, theta, batch, side, wv_frozen=None, label="", atol=ATOL):
    c_ref, g_ref, aux = ntn_ref.compute_at(theta, p, *batch, side, wv_frozen=wv_frozen, return_aux=True)
    assert aux["min_abs_margin"] > 2e-5, f"{label}: oracle sees a near-tie, pick another seed"
    cost, grad, nact = h.eval(theta, side)
    assert nact == aux["n_active"], label
    assert abs(cost - c_ref) <= ATOL + RTOL * abs(c_ref), (label, cost, c_ref)
    err = np.abs(grad - g_ref) - (atol + RTOL * np.abs(g_ref))
    assert err.max() <= 0, (label, int(err.argmax()), float(grad[err.argmax()]), float(g_ref[err.argmax()]))
    # tighter, norm-wise: per parameter block relative L2 error
    offs = list(p.offsets()) + [p.P]
    for a, b in zip(offs[:-1], offs[1:]):
        nrm = np.linalg.norm(g_ref[a:b])
        if nrm > 0:
            assert np.linalg.norm(grad[a:b] - g_ref[a:b]) <= 2e-5 * nrm, (label, a)
    return cost, grad, aux


@pytest.mark.parametrize("gemm_path", ["simt", "auto"])
@pytest.mark.parametrize("shape", [
    dict(Ne=300, R=4, T=500, Nw=200, d=20, k=3, batch=400, corrupt=5),       # everyday small case
    dict(Ne=40, R=3, T=60, Nw=30, d=6, k=3, batch=25, corrupt=3),            # README diagram size (d=6)
    dict(Ne=500, R=11, T=3000, Nw=400, d=100, k=3, batch=2000, corrupt=10),  # reference d,k,C
    dict(Ne=200, R=2, T=300, Nw=100, d=37, k=1, batch=200, corrupt=2),       # odd d, one slice
    # the two stress shapes (k*d+k >= 1064-term contractions) are run at the reference's scale: the divisor of cost and
    # gradient is Run_NTN's batchSize = 20000 (Run_NTN.java:72, NTN.java:398-401), not the 260-300 triples of the toy job
    # -- with a toy divisor the entries are ~67x larger than anything the reference produces and 1e-6 ABSOLUTE is below
    # what FP32 partial sums resolve on either contraction path.  No tolerance is relaxed anywhere in this file.
    dict(Ne=150, R=5, T=400, Nw=80, d=132, k=8, batch=300, corrupt=3, divisor=20000),       # max k, slices split over lanes
    dict(Ne=120, R=3, T=300, Nw=60, d=256, k=8, batch=260, corrupt=2, divisor=20000),       # the scaled config's d and k (BASELINE configs[4])
])
def test_cost_and_gradient_match_oracle(shape, gemm_path):
    kb, p, batch, side, theta0 = small_problem(**shape, seed=1)
    theta = f32(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0))
    h = make_handle(p, "theta", gemm_path)
    h.set_batch(*batch)
    atol = ATOL
    _, _, aux = check(h, p, theta, batch, side, label=str(shape), atol=atol)
    assert 0 < aux["n_active"] < len(batch[0])
    # the other corrupt side for every relation, same batch (NTN.java:146-156 both branches)
    check(h, p, theta, batch, 1 - side, label="flipped sides", atol=atol)
    h.close()


def test_initial_theta_nearly_all_columns_active():
    """theta0 of NTN.java:67-74: scores are close together, nearly every column is inside the margin."""
    kb, p, batch, side, theta0 = small_problem(Ne=300, R=4, T=500, Nw=200, d=20, batch=400, corrupt=5, seed=2)
    h = make_handle(p, "theta")
    h.set_batch(*batch)
    _, _, aux = check(h, p, f32(theta0), batch, side)
    assert aux["n_active"] > 0.9 * len(batch[0])
    h.close()


def test_frozen_word_vectors_reference_behaviour():
    """App. B1: forward built from the load-time vectors while theta's word block drifts."""
    kb, p, batch, side, theta0 = small_problem(Ne=300, R=4, T=500, Nw=200, d=20, batch=400, corrupt=5, seed=3)
    theta = f32(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=1.0))
    h = make_handle(p, "frozen", kb=kb)
    h.set_batch(*batch)
    check(h, p, theta, batch, side, wv_frozen=kb.wv)
    h.close()


def test_sigmoid_activation():
    kb, p, batch, side, theta0 = small_problem(Ne=200, R=3, T=400, Nw=120, d=24, batch=300, corrupt=4, seed=4,
                                               activation="sigmoid")
    theta = f32(spread(perturbed(theta0, 0.05), p))
    h = make_handle(p, "theta")
    h.set_batch(*batch)
    check(h, p, theta, batch, side)
    h.close()


def test_ragged_inputs_empty_relations_duplicates_and_hubs():
    """Edge cases: relations with no column, duplicate training triples (a unique triple with 2C
    corrupt copies), a unique triple whose copies are split irregularly, a hub entity appearing in
    hundreds of columns, and a hub word shared by most entities (the "unknown" fallback)."""
    kb, p, batch, side, theta0 = small_problem(Ne=1200, R=6, T=900, Nw=60, d=16, batch=700, corrupt=3, seed=5,
                                               oov_frac=0.4)
    e1, rel, e2, e3 = (x.copy() for x in batch)
    rel[rel == 2] = 0                                   # relation 2 empty
    e1[::7] = e1[0]; e2[::7] = e2[0]; rel[::7] = rel[0]  # many duplicates of one triple
    e3[::2] = 5                                          # hub corrupt entity (> NTN_HUB_SEG incidents)
    e1[1::3] = 9                                         # hub e1
    keep = np.ones(len(e1), bool)
    keep[5::11] = False                                  # ragged: some triples lose some copies
    b2 = tuple(x[keep] for x in (e1, rel, e2, e3))
    theta = f32(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0))
    h = make_handle(p, "theta")
    h.set_batch(*b2)
    side2 = data_ref.draw_corrupt_sides(b2[1], p.R, JavaRandom(99))
    check(h, p, theta, b2, side2, label="ragged")
    # column order must not matter (up to the FP32 summation order inside a triple's corrupt list)
    perm = np.random.default_rng(0).permutation(len(b2[0]))
    h.set_batch(*(x[perm] for x in b2))
    c1, g1, _ = h.eval(theta, side2)
    h.set_batch(*b2)
    c2, g2, _ = h.eval(theta, side2)
    assert c1 == pytest.approx(c2, rel=1e-6) and np.allclose(g1, g2, rtol=1e-5, atol=1e-8)
    h.close()


def test_empty_batch_is_pure_regulariser():
    kb, p, batch, side, theta0 = small_problem()
    h = make_handle(p, "theta")
    z = np.zeros(0, np.int32)
    h.set_batch(z, z, z, z)
    theta = f32(theta0)
    cost, grad, nact = h.eval(theta, side)
    assert nact == 0
    assert cost == pytest.approx(0.5 * p.lam * np.sum(theta ** 2), rel=1e-6)
    assert np.allclose(grad, p.lam * theta, rtol=1e-6, atol=0)
    h.close()


def test_malformed_entity_names_forward_backward_lists_differ():
    """App. B8 quirk: "__cat__1" averages 1 word forward but spreads its gradient over 2 slots
    (the second is word 0) divided by 2 -- the backward CSR is a separate input."""
    kb, p, batch, side, theta0 = small_problem(Ne=60, R=2, T=100, Nw=40, d=8, batch=60, corrupt=2, seed=6)
    names = list(kb.entity_names)
    names[3] = "__w1__1"
    names[7] = "__w2__4"
    vocab = {w: i for i, w in enumerate(kb.vocab_words)}
    fo, fi, bo, bi, lb = data_ref.build_entity_word_maps(names, vocab)
    assert not np.array_equal(fo, bo)
    p.fwd_off, p.fwd_idx, p.bwd_off, p.bwd_idx, p.len_bwd = fo, fi, bo, bi, lb
    e1, rel, e2, e3 = (x.copy() for x in batch)
    e1[:10] = 3
    e3[:10] = 7
    b2 = (e1, rel, e2, e3)
    theta = f32(spread(perturbed(theta0, 0.05), p))
    h = make_handle(p, "theta")
    h.set_batch(*b2)
    check(h, p, theta, b2, side)
    h.close()


def test_bit_reproducible_across_calls():
    """No float atomics anywhere: two evaluations of the same inputs give identical bits."""
    kb, p, batch, side, theta0 = small_problem(Ne=500, R=11, T=3000, Nw=400, d=100, batch=2000, corrupt=10)
    theta = f32(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0))
    h = make_handle(p, "theta")
    h.set_batch(*batch)
    c1, g1, n1 = h.eval(theta, side)
    c2, g2, n2 = h.eval(theta, side)
    assert c1 == c2 and n1 == n2 and np.array_equal(g1, g2)
    h.close()


def test_float32_entry_point_and_device_resident_path():
    import torch
    kb, p, batch, side, theta0 = small_problem(Ne=300, R=4, T=500, Nw=200, d=20, batch=400, corrupt=5, seed=7)
    theta = f32(spread(perturbed(theta0, 0.05), p))
    h = make_handle(p, "theta")
    h.set_batch(*batch)
    c64, g64, n64 = h.eval(theta, side)
    c32, g32, n32 = h.eval_f32(theta.astype(np.float32), side)
    assert c32 == c64 and n32 == n64 and np.array_equal(g32.astype(np.float64), g64)
    # device-resident: internal layout round trip + eval_dev on torch-owned buffers
    t_ref = torch.from_numpy(theta.astype(np.float32)).cuda()
    t_int = torch.empty(h.Pint, dtype=torch.float32, device="cuda")
    g_int = torch.empty_like(t_int)
    g_ref = torch.empty_like(t_ref)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    h.to_internal_dev(t_ref.data_ptr(), t_int.data_ptr())
    cost, nact = h.eval_dev(t_int.data_ptr(), side, g_int.data_ptr())
    h.from_internal_dev(g_int.data_ptr(), g_ref.data_ptr())
    torch.cuda.synchronize()
    assert cost == c64 and nact == n64
    assert np.array_equal(g_ref.cpu().numpy().astype(np.float64), g64)
    back = torch.empty_like(t_ref)
    h.from_internal_dev(t_int.data_ptr(), back.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(back, t_ref)
    assert np.array_equal(h.download_theta(t_int.data_ptr()), theta)
    h.reset_stream()
    h.close()


def test_reg_scale_makes_rank_results_additive():
    """Data-parallel decomposition: G handles with reg_scale = 1/G on disjoint column
    shards sum to the single-handle result."""
    kb, p, batch, side, theta0 = small_problem(Ne=300, R=4, T=500, Nw=200, d=20, batch=400, corrupt=5, seed=8)
    theta = f32(spread(perturbed(theta0, 0.05), p))
    h = make_handle(p, "theta")
    h.set_batch(*batch)
    c, g, n = h.eval(theta, side)
    G = 4
    acc_c, acc_g, acc_n = 0.0, np.zeros_like(g), 0
    n_cols = len(batch[0])
    for rank in range(G):
        sel = (np.arange(n_cols) % p.batch_size) % G == rank      # shard unique triples, keep their copies together
        hr = make_handle(p, "theta", reg_scale=1.0 / G)
        hr.set_batch(*(x[sel] for x in batch))
        cr, gr, nr = hr.eval(theta, side)
        acc_c += cr; acc_g += gr; acc_n += nr
        hr.close()
    assert acc_n == n
    assert acc_c == pytest.approx(c, rel=1e-6)
    assert np.allclose(acc_g, g, rtol=1e-4, atol=1e-7)
    h.close()


def test_eval_scores_match_oracle():
    kb, p, batch, side, theta0 = small_problem(Ne=300, R=4, T=500, Nw=200, d=20, batch=400, corrupt=5, seed=9)
    theta = f32(spread(perturbed(theta0, 0.05), p))
    e1, rel, e2, _ = batch
    h = make_handle(p, "theta")
    s = h.score(theta, e1[:200], rel[:200], e2[:200], "reference_eval")
    ref = ntn_ref.score_triples_reference_eval(theta, p, e1[:200], rel[:200], e2[:200])
    assert np.allclose(s, ref, rtol=1e-4, atol=1e-5)
    h.close()


def test_errors_are_reported_not_fatal():
    kb, p, batch, side, theta0 = small_problem()
    h = _capi.Handle(p.d, p.k, p.R, p.Ne, p.Nw, p.batch_size, p.lam, wv_source="frozen")
    with pytest.raises(_capi.NTNError) as ei:                     # call order
        h.set_batch(*batch)
        h.eval(theta0, side)
    assert ei.value.code == _capi.NTN_ESTATE
    h.set_entity_words(p.fwd_off, p.fwd_idx, p.bwd_off, p.bwd_idx, p.len_bwd)
    with pytest.raises(_capi.NTNError) as ei:                     # frozen vectors missing
        h.eval(theta0, side)
    assert ei.value.code == _capi.NTN_ESTATE
    bad = batch[3].copy(); bad[0] = p.Ne
    with pytest.raises(_capi.NTNError) as ei:
        h.set_batch(batch[0], batch[1], batch[2], bad)
    assert ei.value.code == _capi.NTN_EINVAL
    lb = p.len_bwd.copy(); lb[2] = -1
    with pytest.raises(_capi.NTNError) as ei:
        h.set_entity_words(p.fwd_off, p.fwd_idx, p.bwd_off, p.bwd_idx, lb)
    assert ei.value.code == _capi.NTN_EINVAL
    h.close()


def test_entity_with_zero_length_name_loads_and_is_skipped_in_the_word_gradient():
    """The real Wordnet entity list holds the name "__2" (DataFactory.java:650 comments on it): entityLength = 0, where
    the reference divides that entity's gradient column by zero (NTN.java:420, App. B8).  The library accepts it and gives
    the entity weight 0 in the word gradient; everything else equals the oracle run in which that entity has an empty
    backward list."""
    kb, p, batch, side, theta0 = small_problem(Ne=200, R=3, T=300, Nw=100, d=12, k=3, batch=200, corrupt=4, seed=17)
    names = list(kb.entity_names)
    victim = int(batch[3][0])                                        # a corrupt entity of the batch: its gradient column is non-zero
    names[victim] = "__2"
    vocab = {w: i for i, w in enumerate(kb.vocab_words)}
    fo, fi, bo, bi, lb = data_ref.build_entity_word_maps(names, vocab)
    assert lb[victim] == 0 and fo[victim + 1] - fo[victim] == 1     # forward: the "unknown" fallback
    theta = f32(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0))
    h = _capi.Handle(p.d, p.k, p.R, p.Ne, p.Nw, p.batch_size, p.lam, wv_source="theta")
    h.set_entity_words(fo, fi, bo, bi, lb)
    h.set_batch(*batch)
    cost, grad, nact = h.eval(theta, side)
    h.close()
    assert np.all(np.isfinite(grad))
    # oracle: same forward maps, the victim's backward list removed (and a harmless divisor)
    keep = np.ones(len(bi), bool); keep[bo[victim]:bo[victim + 1]] = False
    bo2 = bo.copy(); bo2[victim + 1:] -= (bo[victim + 1] - bo[victim])
    lb2 = lb.copy(); lb2[victim] = 1
    p2 = ntn_ref.NTNProblem(d=p.d, k=p.k, R=p.R, Ne=p.Ne, Nw=p.Nw, lam=p.lam, batch_size=p.batch_size,
                            fwd_off=fo, fwd_idx=fi, bwd_off=bo2, bwd_idx=bi[keep], len_bwd=lb2)
    c_ref, g_ref, aux = ntn_ref.compute_at(theta, p2, *batch, side, return_aux=True)
    assert aux["min_abs_margin"] > 2e-5 and nact == aux["n_active"]
    assert abs(cost - c_ref) <= ATOL + RTOL * abs(c_ref)
    assert np.all(np.abs(grad - g_ref) <= ATOL + RTOL * np.abs(g_ref))


@pytest.mark.parametrize("shape", [
    dict(Ne=500, R=11, T=3000, Nw=400, d=100, k=3, batch=2000, corrupt=10),
    dict(Ne=300, R=4, T=500, Nw=200, d=20, k=3, batch=400, corrupt=5),
    dict(Ne=150, R=5, T=400, Nw=80, d=132, k=8, batch=300, corrupt=3),
])
def test_tcgen05_contractions_match_simt_intermediates(shape):
    """The three tcgen05/TMEM 3xTF32 contractions against the FP32 CUDA-core kernels on the same
    inputs: T' (forward), GA (anchor-entity gradient), dWaug split-K partials."""
    kb, p, batch, side, theta0 = small_problem(**shape, seed=1)
    theta = f32(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0))
    out = {}
    for path in ("simt", "tcgen05"):
        h = make_handle(p, "theta", path)
        h.set_batch(*batch)
        h.eval(theta, side)
        assert h.gemm_path == path
        out[path] = {n: h.debug_fetch(n) for n in ("T", "M", "GA", "dWpart")}
        h.close()
    for name in ("T", "M", "GA", "dWpart"):
        a, b = out["simt"][name], out["tcgen05"][name]
        scale = np.abs(a).max()
        # T' is a direct product; M, GA, dWpart are downstream of tanh / the hinge coefficients
        tol = 2e-6 if name == "T" else 3e-5
        assert np.abs(a - b).max() <= tol * scale + 1e-7, (name, float(np.abs(a - b).max()), float(scale))


def test_full_size_wordnet_shape_against_oracle():
    """BASELINE.json configs[1]: Wordnet-shaped, d=100, k=3, 20,000 triples x 10 corrupt, default
    (tcgen05) contraction path, against the float64 oracle at the north-star tolerance.  The oracle
    needs ~15 s of CPU here."""
    from neural_tensor_network_for_kb_completion_b200 import synth
    kb = synth.make_shaped("wordnet", d=100, Nw=67447, T=20000, seed=0)
    vocab = {w: i for i, w in enumerate(kb.vocab_words)}
    fo, fi, bo, bi, lb = data_ref.build_entity_word_maps(kb.entity_names, vocab)
    p = ntn_ref.NTNProblem(d=100, k=3, R=kb.R, Ne=kb.Ne, Nw=kb.Nw, lam=1e-4, batch_size=20000,
                           fwd_off=fo, fwd_idx=fi, bwd_off=bo, bwd_idx=bi, len_bwd=lb)
    batch = data_ref.generate_batch_job(kb.train[:, 0], kb.train[:, 1], kb.train[:, 2], 20000, 10, kb.Ne, JavaRandom(42))
    side = data_ref.draw_corrupt_sides(batch[1], kb.R, JavaRandom(7))
    theta = f32(perturbed(ntn_ref.initial_theta(p, kb.wv, dtype=np.float64), 0.05))
    c_ref, g_ref, aux = ntn_ref.compute_at(theta, p, *batch, side, return_aux=True)
    h = make_handle(p, "theta", "auto")
    h.set_batch(*batch)
    cost, grad, nact = h.eval(theta, side)
    assert h.gemm_path == "tcgen05"
    # a column whose margin is within FP32 noise of the hinge may legitimately flip; each flip moves
    # n_active by one and a handful of gradient entries by O(1/B) = 5e-5
    flips = abs(nact - aux["n_active"])
    assert flips <= 2, (nact, aux["n_active"], aux["min_abs_margin"])
    assert abs(cost - c_ref) <= ATOL + RTOL * abs(c_ref), (cost, c_ref)
    bad = np.abs(grad - g_ref) > (ATOL + RTOL * np.abs(g_ref))
    assert bad.sum() <= 400 * flips, (int(bad.sum()), flips, float(np.abs(grad - g_ref).max()))
    offs = list(p.offsets()) + [p.P]
    for a, b in zip(offs[:-1], offs[1:]):
        nrm = np.linalg.norm(g_ref[a:b])
        assert np.linalg.norm(grad[a:b] - g_ref[a:b]) <= (2e-5 + 1e-3 * flips) * nrm, a
    print(f"full-size: cost {cost:.8f} vs {c_ref:.8f}; n_active {nact} vs {aux['n_active']}; "
          f"max|dgrad| {np.abs(grad - g_ref).max():.3e}; min|margin| {aux['min_abs_margin']:.2e}")
    h.close()


def test_device_lbfgs_matches_numpy_definition():
    """csrc/lbfgs.cu against oracle/lbfgs_ref.py (same algorithm, float64, oracle cost/grad) with the corrupt
    sides held fixed so the objective is deterministic: same iteration/evaluation counts, same decrease."""
    from neural_tensor_network_for_kb_completion_b200.lbfgs import LBFGSMinimizer, Opts
    from oracle import lbfgs_ref
    kb, p, batch, side, theta0 = small_problem(Ne=300, R=4, T=500, Nw=200, d=20, batch=400, corrupt=5, seed=12, lam=1e-3)
    theta = f32(theta0)
    h = make_handle(p, "theta")
    h.set_batch(*batch)

    class Fn:                                   # minimal stand-in for the NTN mirror: fixed sides
        handle, device, update = h, 0, 0
        def _sync_batch(self): pass
        def drawCorruptSides(self): return side
    res = LBFGSMinimizer().minimize(Fn(), theta, Opts(maxIters=5))
    ref_theta, ref_f, ref_it, ref_ev, _ = lbfgs_ref.minimize(lambda t: ntn_ref.compute_at(t, p, *batch, side), theta, 5)
    c0 = ntn_ref.compute_at(theta, p, *batch, side)[0]
    assert res.firstObjVal == pytest.approx(c0, rel=1e-5)
    assert res.minObjVal < c0 and ref_f < c0
    assert (res.iters, res.evals) == (ref_it, ref_ev)
    assert res.minObjVal == pytest.approx(ref_f, rel=2e-3)
    assert np.linalg.norm(res.minArg - ref_theta) <= 2e-2 * np.linalg.norm(ref_theta - theta)
    # the returned argument really has the returned objective value
    assert ntn_ref.compute_at(res.minArg, p, *batch, side)[0] == pytest.approx(res.minObjVal, rel=1e-4)
    h.close()


_TABLES = ("u_e1", "u_e2", "u_rel", "c_off", "c_e3", "inc_off", "inc", "eh_first", "eh_count", "eh_begin", "eh_end", "ent_order")


def _ragged_batch(seed):
    kb, p, batch, side, theta0 = small_problem(Ne=1200, R=6, T=900, Nw=60, d=16, batch=700, corrupt=3, seed=seed, oov_frac=0.4)
    e1, rel, e2, e3 = (x.copy() for x in batch)
    rel[rel == 2] = 0
    e1[::7] = e1[0]; e2[::7] = e2[0]; rel[::7] = rel[0]
    e3[::2] = 5
    e1[1::3] = 9
    keep = np.ones(len(e1), bool)
    keep[5::11] = False
    return kb, p, tuple(x[keep] for x in (e1, rel, e2, e3)), side, theta0


@pytest.mark.parametrize("case", ["small", "ragged", "single_column", "one_relation", "wordnet_like"])
def test_device_batch_index_equals_host_builder(case, monkeypatch):
    """batch_index.cu (the default: stable radix sorts on the device) against the host builder kept for this purpose
    (NTN_SET_BATCH_HOST=1): every integer table identical, hence bit-identical cost and gradient.  Replaces the
    per-evaluation relation filtering of DataFactory.java:89-98 / :614-645."""
    if case == "small":
        kb, p, batch, side, theta0 = small_problem(seed=3)
    elif case == "ragged":
        kb, p, batch, side, theta0 = _ragged_batch(5)
    elif case == "single_column":
        kb, p, batch, side, theta0 = small_problem(seed=4)
        batch = tuple(x[:1] for x in batch)
    elif case == "one_relation":
        kb, p, batch, side, theta0 = small_problem(Ne=300, R=1, T=400, Nw=100, d=8, batch=300, corrupt=4, seed=6)
    else:
        kb, p, batch, side, theta0 = small_problem(Ne=5000, R=11, T=6000, Nw=3000, d=20, batch=5000, corrupt=10, seed=8)
    theta = f32(perturbed(theta0, 0.05))
    out = {}
    for mode in ("host", "device"):
        if mode == "host":
            monkeypatch.setenv("NTN_SET_BATCH_HOST", "1")
        else:
            monkeypatch.delenv("NTN_SET_BATCH_HOST", raising=False)
        h = make_handle(p, "theta")
        h.set_batch(*batch)
        tables = {t: h.debug_fetch("i:" + t).copy() for t in _TABLES}
        out[mode] = (tables, h.eval(theta, side))
        h.close()
    for t in _TABLES:
        np.testing.assert_array_equal(out["host"][0][t], out["device"][0][t], err_msg=t)
    (c_h, g_h, n_h), (c_d, g_d, n_d) = out["host"][1], out["device"][1]
    assert c_h == c_d and n_h == n_d and np.array_equal(g_h, g_d)


def test_set_batch_from_device_columns():
    """ntn_set_batch_dev: the batch columns already live in HBM (torch tensors here)."""
    import torch
    kb, p, batch, side, theta0 = _ragged_batch(7)
    theta = f32(perturbed(theta0, 0.05))
    h = make_handle(p, "theta")
    h.set_batch(*batch)
    ref = h.eval(theta, side)
    cols = [torch.from_numpy(np.ascontiguousarray(x, np.int32)).cuda() for x in batch]
    torch.cuda.synchronize()
    h2 = make_handle(p, "theta")
    h2.set_batch_dev(len(batch[0]), *(c.data_ptr() for c in cols))
    got = h2.eval(theta, side)
    assert ref[0] == got[0] and ref[2] == got[2] and np.array_equal(ref[1], got[1])
    # an out-of-range index is found on the device and reported with its column
    cols[3][17] = p.Ne
    torch.cuda.synchronize()
    with pytest.raises(_capi.NTNError) as ei:
        h2.set_batch_dev(len(batch[0]), *(c.data_ptr() for c in cols))
    assert ei.value.code == _capi.NTN_EINVAL and "column 17" in str(ei.value)
    h.close(); h2.close()


@pytest.mark.parametrize("filter_true", [False, True])
def test_device_batch_job_generator(filter_true):
    """ntn_generate_batch_dev (sampler variants): the job generated in HBM equals the NumPy restatement
    column for column, and its evaluation matches the oracle on those columns.  A tiny entity set makes true-triple
    collisions frequent so the filtered redraws are exercised."""
    kb, p, _, side, theta0 = small_problem(Ne=60, R=3, T=900, Nw=40, d=8, batch=300, corrupt=6, seed=9)
    first, B, C, seed, job = 100, 300, 6, 2024, 5
    want = data_ref.generate_batch_job_device(kb.train, first, B, C, p.Ne, seed, job, filter_true)
    h = make_handle(p, "theta")
    h.set_training_triples(kb.train)
    h.generate_batch_dev(first, B, C, seed, job, filter_true)
    got = h.get_batch_columns(B * C)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    plain = data_ref.generate_batch_job_device(kb.train, first, B, C, p.Ne, seed, job, False)
    true = set(map(tuple, kb.train.tolist()))
    hits = lambda cols: sum(((int(a), int(r), int(c)) in true) or ((int(c), int(r), int(b)) in true) for a, r, b, c in zip(*cols))
    assert hits(plain) > 20                                           # the unfiltered job does hit true triples ...
    if filter_true:
        assert hits(got) < hits(plain) / 10                           # ... the filtered one (almost) never
    theta = f32(perturbed(theta0, 0.05))
    side = data_ref.draw_corrupt_sides(got[1], p.R, JavaRandom(3))
    check(h, p, theta, got, side, label="device job")
    # a second job: different counter, different columns
    h.generate_batch_dev(first, B, C, seed, job + 1, filter_true)
    assert not np.array_equal(h.get_batch_columns(B * C)[3], got[3])
    with pytest.raises(_capi.NTNError):
        h.generate_batch_dev(700, B, C, seed, job, filter_true)      # range outside the training triples
    h.close()


def test_training_loop_with_device_sampler():
    """NTN.generateBatchJobOnDevice + device L-BFGS: the Run_NTN loop without any host batch job."""
    from neural_tensor_network_for_kb_completion_b200 import synth
    from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory
    from neural_tensor_network_for_kb_completion_b200.lbfgs import LBFGSMinimizer, Opts
    from neural_tensor_network_for_kb_completion_b200.ntn import NTN
    kb = synth.make_structured_kb(Ne=300, R=3, T=1500, d=12, seed=2)
    tbj = DataFactory.from_kb(kb, 500, 4, seed=42)
    t = NTN(kb.d, kb.Ne, kb.R, kb.Nw, 500, 2, "tanh", tbj, 1e-4, wv_source="theta")
    theta = np.asarray(t.getTheta_inital(), np.float64)
    costs = []
    for it in range(4):
        t.generateBatchJobOnDevice(seed=11, job=it, filter_true=True)
        res = LBFGSMinimizer().minimize(t, theta, Opts(maxIters=5))
        theta = res.minArg
        costs.append((res.firstObjVal, res.minObjVal))
    assert all(b <= a for a, b in costs) and costs[-1][1] < costs[0][0]
    cols = t.generateBatchJobOnDevice(seed=11, job=0, filter_true=True, fetch=True)
    want = data_ref.generate_batch_job_device(kb.train, 0, 500, 4, kb.Ne, 11, 0, True)
    for a, b in zip(cols, want):
        np.testing.assert_array_equal(a, b)
    t.close()


def test_full_size_freebase_shape_properties():
    """BASELINE.json configs[2] (the headline: Freebase shape, 75,043 entities, 13 relations, d=100, k=3, 20,000 triples
    x 10 corrupt) is too slow for the float64 oracle in a test, so it is checked through size-independent properties:
      1. bit-reproducibility (no float atomics anywhere);
      2. linearity over shards: 4 handles with reg_scale = 1/4 on disjoint shards sum to the full result
         (the data-parallel decomposition);
      3. the gradient is the derivative of the cost: central difference of the COST along a random direction and along
         the gradient direction against g.v (FP32 cost, so a loose tolerance; hinge kinks only move it slightly);
      4. an empty-relation / regulariser identity: d cost / d theta restricted to words no entity uses equals lambda*theta."""
    from neural_tensor_network_for_kb_completion_b200 import synth
    from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory
    from neural_tensor_network_for_kb_completion_b200.distributed import shard_columns
    kb = synth.make_shaped("freebase", d=100, Nw=67447, T=20000, seed=0)
    tbj = DataFactory.from_kb(kb, 20000, 10, seed=42)
    maps = tbj.entityWordMaps()
    batch = tbj.generateNewTrainingBatchJob()
    lam, B = 1e-4, 20000

    def handle(reg_scale=1.0):
        h = _capi.Handle(100, 3, kb.R, kb.Ne, kb.Nw, B, lam, wv_source="theta", reg_scale=reg_scale)
        h.set_entity_words(*maps)
        return h
    h = handle()
    h.set_batch(*batch)
    assert h.gemm_path == "tcgen05"
    rng = np.random.default_rng(5)
    r = 1.0 / np.sqrt(200.0)
    theta = np.concatenate([rng.random(kb.R * 3 * 100 * 100) * r, np.full(kb.R * 600, 0.4), np.full(kb.R * 3, 0.01),
                            np.full(kb.R * 3, 0.8), kb.wv.astype(np.float64).ravel()])
    theta = f32(theta + rng.standard_normal(theta.size) * 0.05)
    side = data_ref.draw_corrupt_sides(batch[1], kb.R, JavaRandom(7))
    c, g, n = h.eval(theta, side)
    assert 0 < n < 200000 and np.isfinite(c) and np.all(np.isfinite(g))
    # 1. reproducible
    c2, g2, n2 = h.eval(theta, side)
    assert c == c2 and n == n2 and np.array_equal(g, g2)
    # 2. shard linearity
    G = 4
    acc_c, acc_g, acc_n = 0.0, np.zeros_like(g), 0
    for rank in range(G):
        hr = handle(1.0 / G)
        hr.set_batch(*shard_columns(*batch, rank, G))
        cr, gr, nr = hr.eval(theta, side)
        acc_c += cr; acc_g += gr; acc_n += nr
        hr.close()
    assert acc_n == n and acc_c == pytest.approx(c, rel=2e-6)
    assert np.allclose(acc_g, g, rtol=1e-4, atol=2e-7)
    # 3. directional derivatives
    for v in (rng.standard_normal(theta.size), g.copy()):
        v /= np.linalg.norm(v)
        eps = 2e-3
        cp = h.eval(f32(theta + eps * v), side)[0]
        cm = h.eval(f32(theta - eps * v), side)[0]
        th_p, th_m = f32(theta + eps * v), f32(theta - eps * v)
        fd = (cp - cm) / float(np.dot(th_p - th_m, v))
        gv = float(np.dot(g, v))
        assert abs(fd - gv) <= 2e-2 * max(abs(gv), 0.05), (fd, gv)
    # 4. unused words see the regulariser only
    used = np.zeros(kb.Nw, bool)
    used[maps[3]] = True                                           # backward word lists (NTN.java:414-426)
    unused = np.nonzero(~used)[0][:200]
    oWV = theta.size - 100 * kb.Nw
    for w in unused[:50]:
        idx = oWV + np.arange(100) * kb.Nw + w                     # word block is d x Nw row-major (NTN.java:450-481)
        assert np.allclose(g[idx], lam * theta[idx], rtol=1e-5, atol=1e-12)
    h.close()


@pytest.mark.parametrize("shape", [
    dict(Ne=300, R=4, T=500, Nw=200, d=20, k=3, batch=400, corrupt=5),
    dict(Ne=200, R=2, T=300, Nw=100, d=37, k=1, batch=200, corrupt=2),
    dict(Ne=150, R=5, T=400, Nw=80, d=132, k=8, batch=300, corrupt=3, divisor=20000),
])
def test_large_batch_gv_variant_matches_oracle(shape, monkeypatch):
    """The large-batch variant (k_score materialises one contribution row per incident, picked automatically when T'
    outgrows the L2) forced on small problems: same oracle, same tolerance."""
    monkeypatch.setenv("NTN_GV", "1")
    kb, p, batch, side, theta0 = small_problem(**shape, seed=21)
    theta = f32(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0))
    h = make_handle(p, "theta", "simt")
    h.set_batch(*batch)
    check(h, p, theta, batch, side, label="gv " + str(shape))
    check(h, p, theta, batch, 1 - side, label="gv flipped")
    h.close()


@pytest.mark.parametrize("corrupt", [1, 10, 11, 23])
def test_corrupt_list_lengths_around_the_score_group(corrupt):
    """k_score2 handles a unique triple's corrupt columns in groups of 10 (corrupt_size of Run_NTN.java:77): shorter,
    exact, one more and several groups (23 = duplicates of a training triple merged into one unique triple)."""
    kb, p, batch, side, theta0 = small_problem(Ne=300, R=3, T=200, Nw=150, d=12, k=3, batch=120, corrupt=corrupt, seed=30 + corrupt)
    theta = f32(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0))
    for path in ("simt", "auto"):
        h = make_handle(p, "theta", path)
        h.set_batch(*batch)
        check(h, p, theta, batch, side, label=f"corrupt={corrupt} {path}")
        h.close()


@pytest.mark.parametrize("d", [6, 37])
def test_internal_padding_columns_stay_zero(d):
    """d % 4 != 0: the internal word rows are ceil4(d) wide.  The padding columns are not parameters: their gradient
    must be exactly 0 (the anchor rows of the entity-gradient contraction carry the bias product at column d) and the
    device L-BFGS must leave theta's padding at 0."""
    import torch
    kb, p, batch, side, theta0 = small_problem(Ne=200, R=3, T=300, Nw=100, d=d, k=3, batch=200, corrupt=4, seed=40 + d)
    theta = f32(spread(perturbed(theta0, 0.05), p, u=4.0, w=10.0, wv=3.0))
    dq = (d + 3) // 4 * 4
    for path in ("simt", "auto"):
        h = make_handle(p, "theta", path)
        h.set_batch(*batch)
        d_theta = torch.zeros(h.Pint, dtype=torch.float32, device="cuda")
        d_grad = torch.full((h.Pint,), 7.0, dtype=torch.float32, device="cuda")
        h.upload_theta(theta, d_theta.data_ptr())
        torch.cuda.synchronize()
        h.eval_dev(d_theta.data_ptr(), side, d_grad.data_ptr())
        o_wv = h.Pint - p.Nw * dq
        g = d_grad[o_wv:].reshape(p.Nw, dq).cpu().numpy()
        t = d_theta[o_wv:].reshape(p.Nw, dq).cpu().numpy()
        assert np.all(g[:, d:] == 0.0), (path, float(np.abs(g[:, d:]).max()))
        assert np.all(t[:, d:] == 0.0)
        assert np.abs(g[:, :d]).max() > 0
        h.lbfgs_minimize(d_theta.data_ptr(), lambda: side, max_iters=3)
        t = d_theta[o_wv:].reshape(p.Nw, dq).cpu().numpy()
        assert np.all(t[:, d:] == 0.0), path
        h.close()


def test_headline_config_full_size_against_golden():
    """BASELINE.json configs[2] at FULL size -- the exact inputs bench.py times (75,043 entities, 13 relations, d=100,
    k=3, 20,000 x 10 columns, tcgen05 contractions) -- against the committed float64-oracle golden vector
    (tests/golden/c3_fullsize.npz; generator tests/golden/make_c3_golden.py) at the north-star tolerance
    1e-6 + 1e-4|ref| on cost and every stored gradient entry (1 in 37 of the 7.1 M, plus all of V, b, U), n_active exact.
    The generator picked theta's seed so that the oracle sees no hinge margin within 2e-5 of zero (same near-tie guard as
    check()), hence no flip allowance here."""
    import bench
    g = np.load(bench.GOLDEN_C3)
    assert float(g["min_abs_margin"]) > 2e-5
    kb, p, maps, batch, side, theta0 = bench.headline_problem()
    theta = bench.theta_trained_like(theta0, seed=int(g["theta_seed"])).astype(np.float64)
    h = make_handle(p, "theta", "auto")
    h.set_batch(*batch)
    cost, grad, nact = h.eval(theta, side)
    assert h.gemm_path == "tcgen05"
    res = bench.check_against_golden(cost, grad, nact, theta)
    print(res)
    assert res["status"] == "pass", res
    assert res["block_l2_rel_diff_max"] <= 2e-5
    # the float32 host entry point returns the same numbers
    c32, g32, n32 = h.eval_f32(theta.astype(np.float32), side)
    assert n32 == nact and c32 == cost and np.array_equal(g32.astype(np.float64), grad)
    h.close()
