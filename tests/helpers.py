"""Shared builders for the tests: a small synthetic KB turned into oracle inputs."""
import numpy as np

from neural_tensor_network_for_kb_completion_b200 import synth
from oracle import data_ref, ntn_ref
from oracle.java_random import JavaRandom


def small_problem(Ne=40, R=3, T=60, Nw=30, d=6, k=3, batch=25, corrupt=3, lam=1e-4,
                  seed=0, activation="tanh", oov_frac=0.05, divisor=None):
    """divisor: the reference's batchSize, the divisor of cost and gradient (NTN.java:398-401,429-430); defaults to
    the number of training triples in the batch job."""
    kb = synth.make_kb(Ne, R, T, Nw, d, oov_frac=oov_frac, seed=seed)
    vocab = {w: i for i, w in enumerate(kb.vocab_words)}
    fo, fi, bo, bi, lb = data_ref.build_entity_word_maps(kb.entity_names, vocab)
    p = ntn_ref.NTNProblem(d=d, k=k, R=R, Ne=Ne, Nw=Nw, lam=lam, batch_size=divisor or batch,
                           fwd_off=fo, fwd_idx=fi, bwd_off=bo, bwd_idx=bi, len_bwd=lb,
                           activation=activation)
    e1, rel, e2, e3 = data_ref.generate_batch_job(kb.train[:, 0], kb.train[:, 1], kb.train[:, 2],
                                                  batch, corrupt, Ne, JavaRandom(42 + seed))
    side = data_ref.draw_corrupt_sides(rel, R, JavaRandom(7 + seed))
    theta0 = ntn_ref.initial_theta(p, kb.wv, seed=3 + seed, dtype=np.float64)
    return kb, p, (e1, rel, e2, e3), side, theta0


def perturbed(theta0, scale=0.05, seed=11):
    """A 'trained-like' theta: theta0 + N(0, scale^2), so the hinge is
    not active everywhere and U,V,b are not constants."""
    rng = np.random.default_rng(seed)
    return theta0 + rng.standard_normal(theta0.shape) * scale


def spread(theta, p, u=6.0, w=30.0, wv=5.0):
    """Scale U, W and the word block so that scores spread out and the hinge is inactive
    on a good part of the columns (both branches of NTN.java:213-217 exercised)."""
    oW, oV, ob, oU, oWV = p.offsets()
    t = theta.copy()
    t[oU:oWV] *= u
    t[oW:oV] *= w
    t[oWV:] *= wv
    return t
