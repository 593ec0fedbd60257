"""NTN.computeAt restated on the CPU (oracle; test infra -- see oracle/__init__.py).

Follows /root/reference/src/NTN.java:92-443 statement by statement, in the
reference's own *per-relation, per-slice, per-column* formulation -- it does NOT
use the de-duplication identity the CUDA path is built on, so
the two are independent derivations of the same numbers.

PARITY UNPINNED: the reference cannot be executed here (no JVM / jars) and holds
no golden vectors; ND4j's flattening order is frozen to row-major per block
(SURVEY.md App. B9).  What pins this file instead: tests/test_oracle.py (central
finite differences, float32-vs-float64 agreement, hand-computed tiny case).

dtype=np.float64 is the parity oracle; dtype=np.float32 executes the same
statements in the reference's arithmetic type (Run_NTN.java:55) and is what
bench.py times as the "reference-formulation CPU stand-in".
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .data_ref import create_entity_vectors


@dataclass
class NTNProblem:
    """Everything NTN's constructor + the DataFactory singleton hold
    (NTN.java:42-66, DataFactory.java:55-63)."""
    d: int                      # embeddingSize
    k: int                      # sliceSize
    R: int                      # numberOfRelations
    Ne: int                     # numberOfEntities
    Nw: int                     # numOfWords
    lam: float                  # lamda
    batch_size: int             # the divisor (NTN.java:398-401,429-430)
    fwd_off: np.ndarray         # entity -> forward words CSR
    fwd_idx: np.ndarray
    bwd_off: np.ndarray         # entity -> backward words CSR (getWordIndexes)
    bwd_idx: np.ndarray
    len_bwd: np.ndarray         # entityLength(i)
    activation: str = "tanh"    # "tanh" | "sigmoid" (NTN.java:30,752-756)
    extra: dict = field(default_factory=dict)

    @property
    def P(self) -> int:
        d, k, R = self.d, self.k, self.R
        return R * k * d * d + R * 2 * d * k + R * k + R * k + d * self.Nw

    def offsets(self):
        d, k, R = self.d, self.k, self.R
        oW = 0
        oV = oW + R * k * d * d
        ob = oV + R * 2 * d * k
        oU = ob + R * k
        oWV = oU + R * k
        return oW, oV, ob, oU, oWV


def stack_to_parameters(theta: np.ndarray, p: NTNProblem):
    """NTN.java:483-526: W[r] (k,d,d) by linear index, then V[r] (2d,k), b[r] (1,k),
    U[r] (k,1), then word vectors (d, Nw); row-major per block (App. B9)."""
    d, k, R = p.d, p.k, p.R
    oW, oV, ob, oU, oWV = p.offsets()
    W = theta[oW:oV].reshape(R, k, d, d)
    V = theta[oV:ob].reshape(R, 2 * d, k)
    b = theta[ob:oU].reshape(R, k)
    U = theta[oU:oWV].reshape(R, k)
    WV = theta[oWV:oWV + d * p.Nw].reshape(d, p.Nw)
    return W, V, b, U, WV


def parameters_to_stack(W, V, b, U, WV) -> np.ndarray:
    """NTN.java:450-481."""
    return np.concatenate([W.ravel(), V.ravel(), b.ravel(), U.ravel(), WV.ravel()])


def initial_theta(p: NTNProblem, wv_loaded: np.ndarray, seed: int = 3, dtype=np.float32) -> np.ndarray:
    """NTN.java:67-84: W = rand * (2r - r) = rand * r with r = 1/sqrt(2d) (App. B11);
    V = 0.4, b = 0.01, U = 0.8; word block = the loaded vectors.  ND4j's RNG is not
    reproducible here -- values, not RNG bits, are what matters."""
    rng = np.random.default_rng(seed)
    r = 1.0 / np.sqrt(2.0 * p.d)
    W = rng.random((p.R, p.k, p.d, p.d)) * r
    V = np.full((p.R, 2 * p.d, p.k), 0.4)
    b = np.full((p.R, p.k), 0.01)
    U = np.full((p.R, p.k), 0.8)
    return parameters_to_stack(W, V, b, U, np.asarray(wv_loaded, np.float64)).astype(dtype)


def _activation(pre, kind):
    if kind == "tanh":
        return np.tanh(pre)
    if kind == "sigmoid":
        return 1.0 / (1.0 + np.exp(-pre))
    raise ValueError(kind)


def _activation_differential(z, kind):
    """NTN.java:744-759: tanh -> 1 - z*z; sigmoid -> z*(1-z) (comment :746)."""
    if kind == "tanh":
        return 1.0 - z * z
    return z * (1.0 - z)


def _scatter_add_columns(dst: np.ndarray, x: np.ndarray, idx: np.ndarray, fast: bool):
    """Util.java:74-98 MatrixX_mmul_CSRMatrix followed by the dense add
    (NTN.java:372,388): dst[:, idx[j]] += x[:, j] for every column j."""
    if x.shape[1] == 0:
        return
    if fast:
        # one-hot CSR product: the thing the reference's helper is named after
        import scipy.sparse as sp
        n = x.shape[1]
        S = sp.csr_matrix((np.ones(n, x.dtype), (np.arange(n), idx)), shape=(n, dst.shape[1]))
        dst += (S.T @ x.T).T
    else:
        np.add.at(dst, (slice(None), idx), x)


def compute_at(theta, p: NTNProblem, e1, rel, e2, e3, corrupt_side,
               wv_frozen=None, dtype=np.float64, fast_scatter=True, return_aux=False):
    """(cost, grad[, aux]) = NTN.computeAt(theta)   [NTN.java:92-443]

    e1, rel, e2, e3 : the batch job columns in batch order (DataFactory.java:116-136)
    corrupt_side[r] : 1 -> negatives (e1, e3) (NTN.java:146-150), 0 -> (e3, e2) (:151-155)
    wv_frozen       : d x Nw load-time word vectors.  If given, entity vectors are built
                      from it (the reference's actual behaviour, App. B1); if None they are
                      built from theta's word block (the differentiable, intended one).
    """
    dt = dtype
    theta = np.asarray(theta).astype(dt)                      # Util.java:49-55 (FP32 buffer)
    d, k, R, B = p.d, p.k, p.R, p.batch_size
    W, V, b, U, WV = stack_to_parameters(theta, p)            # NTN.java:96
    src = WV if wv_frozen is None else np.asarray(wv_frozen).astype(dt)
    E = create_entity_vectors(src, p.fwd_off, p.fwd_idx, dt)  # NTN.java:103
    gE = np.zeros((d, p.Ne), dt)                              # NTN.java:100
    cost = dt(0)
    n_active = 0
    gW = np.zeros((R, k, d, d), dt)
    gV = np.zeros((R, 2 * d, k), dt)
    gb = np.zeros((R, k), dt)
    gU = np.zeros((R, k), dt)
    min_margin = np.inf

    rel = np.asarray(rel)
    for r in range(R):                                        # NTN.java:114
        cols = np.nonzero(rel == r)[0]                        # DataFactory.java:89-98 (batch order)
        if cols.size == 0:                                    # NTN.java:120 / :402-408
            continue
        i1, i2, i3 = e1[cols], e2[cols], e3[cols]             # :123-126
        E1, E2, E3 = E[:, i1], E[:, i2], E[:, i3]             # :138-143
        if corrupt_side[r]:                                   # :146-150
            E1n, E2n, i1n, i2n = E1, E3, i1, i3
        else:                                                 # :151-155
            E1n, E2n, i1n, i2n = E3, E2, i3, i2
        pre_p = np.zeros((k, cols.size), dt)                  # :159-160
        pre_n = np.zeros((k, cols.size), dt)
        for s in range(k):                                    # :164-179
            Ws = W[r, s]
            pre_p[s] = np.sum(E1 * (Ws @ E2), axis=0)
            pre_n[s] = np.sum(E1n * (Ws @ E2n), axis=0)
        Vt = V[r].T                                           # :182
        pre_p = pre_p + (Vt @ np.vstack([E1, E2]) + b[r][:, None])      # :196
        pre_n = pre_n + (Vt @ np.vstack([E1n, E2n]) + b[r][:, None])    # :197
        z_p = _activation(pre_p, p.activation)                # :200-201
        z_n = _activation(pre_n, p.activation)
        s_p = U[r] @ z_p                                      # :204-205
        s_n = U[r] @ z_n
        margin = s_p + dt(1) - s_n
        indx = (s_p + dt(1)) > s_n                            # :209-218 (strict)
        min_margin = min(min_margin, float(np.min(np.abs(margin))))
        cost = cost + np.sum(np.where(indx, s_p - s_n + dt(1), dt(0)), dtype=dt)   # :221
        act = np.nonzero(indx)[0]                             # :229-239
        n_active += act.size                                  # :222
        if act.size == 0:                                     # :272-277,284-289,390-396
            continue
        zpf, znf = z_p[:, act], z_n[:, act]                   # :241-262
        E1f, E2f, E1nf, E2nf = E1[:, act], E2[:, act], E1n[:, act], E2n[:, act]   # :243-246
        i1f, i2f, i1nf, i2nf = i1[act], i2[act], i1n[act], i2n[act]               # :264-268
        gU[r] = np.sum(zpf - znf, axis=1)                     # :273
        t_p = _activation_differential(zpf, p.activation) * U[r][:, None]        # :280
        t_n = _activation_differential(znf, p.activation) * (-U[r])[:, None]     # :281
        gb[r] = np.sum(t_p + t_n, axis=1)                     # :285
        for s in range(k):                                    # :304
            tp, tn = t_p[s], t_n[s]                           # :309-316
            gW[r, s] = (E1f * tp) @ E2f.T + (E1nf * tn) @ E2nf.T                 # :323,336
            gV[r][:, s] = np.sum(np.vstack([E1f, E2f]) * tp + np.vstack([E1nf, E2nf]) * tn, axis=1)  # :342,349
            v_pos = np.outer(V[r][:, s], tp)                  # :352-365
            v_neg = np.outer(V[r][:, s], tn)
            _scatter_add_columns(gE, v_pos[:d], i1f, fast_scatter)    # :367-372
            _scatter_add_columns(gE, v_pos[d:], i2f, fast_scatter)
            _scatter_add_columns(gE, v_neg[:d], i1nf, fast_scatter)
            _scatter_add_columns(gE, v_neg[d:], i2nf, fast_scatter)
            Ws = W[r, s]
            _scatter_add_columns(gE, (Ws @ E2f) * tp, i1f, fast_scatter)       # :384
            _scatter_add_columns(gE, (Ws.T @ E1f) * tp, i2f, fast_scatter)     # :385
            _scatter_add_columns(gE, (Ws @ E2nf) * tn, i1nf, fast_scatter)     # :386
            _scatter_add_columns(gE, (Ws.T @ E1nf) * tn, i2nf, fast_scatter)   # :387
        gW[r] /= dt(B)                                        # :398-401
        gV[r] /= dt(B)
        gb[r] /= dt(B)
        gU[r] /= dt(B)

    # entity -> word gradients, NTN.java:411-427
    gWV = np.zeros((d, p.Nw), dt)
    nz = np.nonzero(np.sum(gE, axis=0, dtype=dt) != 0)[0]     # :417 (exact-zero skip, App. B12)
    if nz.size:
        g = gE[:, nz] / p.len_bwd[nz].astype(dt)              # :420
        cnt = (p.bwd_off[nz + 1] - p.bwd_off[nz])
        src_col = np.repeat(np.arange(nz.size), cnt)
        starts = np.repeat(p.bwd_off[nz], cnt)
        within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        words = p.bwd_idx[starts + within]
        _scatter_add_columns(gWV, g[:, src_col], words, fast_scatter)   # :422-425
    gWV = gWV / dt(B)                                         # :429
    cost = cost / dt(B)                                       # :430
    grad = parameters_to_stack(gW, gV, gb, gU, gWV)           # :433
    cost = cost + dt(0.5) * (dt(p.lam) * np.sum(theta * theta, dtype=dt))   # :437
    grad = grad + theta * dt(p.lam)                           # :438
    if return_aux:
        return float(cost), grad.astype(np.float64), {
            "n_active": int(n_active), "min_abs_margin": float(min_margin), "E": E, "gE": gE}
    return float(cost), grad.astype(np.float64)               # :442


def score_triples_reference_eval(theta, p: NTNProblem, e1, rel, e2, wv_frozen=None, dtype=np.float64):
    """NTN.calculateScoreOfaTripple (NTN.java:713-736): NO activation, U scales only
    the bilinear term:  score = sum_s ( U[s] * (e1' W_s e2) + V[:,s]' [e1;e2] + b[s] )
    (SURVEY.md App. B6)."""
    theta = np.asarray(theta).astype(dtype)
    W, V, b, U, WV = stack_to_parameters(theta, p)
    src = WV if wv_frozen is None else np.asarray(wv_frozen).astype(dtype)
    E = create_entity_vectors(src, p.fwd_off, p.fwd_idx, dtype)
    out = np.zeros(len(e1), np.float64)
    for j in range(len(e1)):
        a, c, r = E[:, e1[j]], E[:, e2[j]], rel[j]
        st = np.concatenate([a, c])                           # :721
        sc = 0.0
        for s in range(p.k):                                  # :726-732
            dot2 = a @ (W[r, s] @ c)
            dot3 = V[r][:, s] @ st
            sc += float(U[r, s] * dot2 + dot3 + b[r, s])
        out[j] = sc
    return out


def compute_best_thresholds(scores, rel, label, R: int):
    """NTN.computeBestThresholds sweep (NTN.java:607-666) with its quirks (App. B7):
    thresholds start at -1 (:619), accuracies at 0; ``score_temp += 0.01`` sits INSIDE
    the relation loop (:662) so relation i only sees candidates min + 0.01*(i + R*m);
    strict '>' on accuracy (:658); a relation with no dev triples gives 0/0 = NaN,
    never '>' the stored accuracy."""
    scores = np.asarray(scores, np.float64)
    rel = np.asarray(rel)
    label = np.asarray(label)
    best_thr = np.full(R, -1.0)
    best_acc = np.zeros(R)
    smin, smax = float(scores.min()), float(scores.max())     # :608-609
    per_rel = [np.nonzero(rel == i)[0] for i in range(R)]
    t = smin                                                  # :622
    while t <= smax:                                          # :627
        for i in range(R):                                    # :629
            idx = per_rel[i]
            if idx.size:
                pred_true = scores[idx] <= t                  # :637
                acc = (np.sum(pred_true & (label[idx] == 1)) +
                       np.sum(~pred_true & (label[idx] == -1))) / idx.size   # :640-654
                if acc > best_acc[i]:                         # :658
                    best_acc[i] = acc
                    best_thr[i] = t
            t = t + 0.01                                      # :662
    return best_thr, best_acc


def get_prediction(scores, rel, label, thresholds):
    """NTN.getPrediction (NTN.java:672-712): +1 iff score <= threshold[rel];
    returns (predictions, mean accuracy)."""
    scores = np.asarray(scores, np.float64)
    pred = np.where(scores <= np.asarray(thresholds)[np.asarray(rel)], 1, -1)
    return pred, float(np.mean(pred == np.asarray(label)))
