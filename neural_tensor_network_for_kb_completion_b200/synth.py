"""Synthetic Wordnet- / Freebase-shaped knowledge bases.

There is no network and the Socher data files are not in the reference tree, so
every test and benchmark runs on data of the reference's *shape*: entity names in
the format DataFactory's parser expects (``__word_word_N``,
/root/reference/src/DataFactory.java:164,399,587), a word-vector matrix d x Nw with
the literal word "unknown" (DataFactory.java:415), ``e1 rel e2`` training triples
and labelled dev/test triples (DataFactory.java:219-325).

Pure NumPy, deterministic in its seeds, and independent of both the CUDA library
and the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SHAPES = {
    # name: (Ne, R, T_train, n_dev, n_test)
    "wordnet": (38696, 11, 112581, 2 * 2609, 2 * 10544),
    "freebase": (75043, 13, 316232, 2 * 5908, 2 * 23733),
}


@dataclass
class SyntheticKB:
    d: int
    entity_names: list            # Ne strings, "__w12_w7_3"
    vocab_words: list             # Nw strings, vocab_words[i] <-> column i of wv
    wv: np.ndarray                # d x Nw float32 (initEmbed.mat "We", one column per word)
    relations: list               # R strings
    train: np.ndarray             # T x 3 int32 (e1, rel, e2)
    dev: np.ndarray               # n x 4 int32 (e1, rel, e2, label in {1,-1})
    test: np.ndarray

    @property
    def Ne(self):
        return len(self.entity_names)

    @property
    def Nw(self):
        return len(self.vocab_words)

    @property
    def R(self):
        return len(self.relations)


def _zipf_choice(rng, n_items: int, a: float, size: int) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), a)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    return np.searchsorted(cdf, rng.random(size), side="right").astype(np.int64)


def make_kb(Ne: int, R: int, T: int, Nw: int, d: int, n_dev: int = 0, n_test: int = 0,
            oov_frac: float = 0.01, seed: int = 0) -> SyntheticKB:
    """Ne entities of 1-4 words (70/22/6/2 %), word ids Zipf(1.1) over the vocabulary,
    a fraction ``oov_frac`` of word slots deliberately out of vocabulary (exercises the
    "unknown" fallback), sense suffix 1..9; WV ~ N(0, 0.1^2) seed+1; triples: rel
    Zipf(1.0), e1/e2 uniform, seed+2."""
    assert Nw >= 2
    rng = np.random.default_rng(seed)
    vocab_words = [f"w{i}" for i in range(Nw - 1)] + ["unknown"]
    lens = rng.choice([1, 2, 3, 4], size=Ne, p=[0.70, 0.22, 0.06, 0.02])
    total = int(lens.sum())
    wid = _zipf_choice(rng, Nw - 1, 1.1, total)
    oov = rng.random(total) < oov_frac
    sense = rng.integers(1, 10, size=Ne)
    names = []
    pos = 0
    for n in range(Ne):
        ws = []
        for j in range(lens[n]):
            ws.append(f"oov{wid[pos]}" if oov[pos] else vocab_words[wid[pos]])
            pos += 1
        names.append("__" + "_".join(ws) + "_" + str(int(sense[n])))
    wv = (np.random.default_rng(seed + 1).standard_normal((d, Nw)) * 0.1).astype(np.float32)
    relations = [f"_rel{i}" for i in range(R)]

    def triples(n, rs):
        if n == 0:
            return np.zeros((0, 3), np.int32)
        rel = _zipf_choice(rs, R, 1.0, n)
        e1 = rs.integers(0, Ne, size=n)
        e2 = rs.integers(0, Ne, size=n)
        return np.stack([e1, rel, e2], axis=1).astype(np.int32)

    rs = np.random.default_rng(seed + 2)
    train = triples(T, rs)

    def labelled(n):
        t = triples(n, rs)
        lab = np.where(np.arange(n) % 2 == 0, 1, -1).astype(np.int32)
        return np.concatenate([t, lab[:, None]], axis=1)

    return SyntheticKB(d=d, entity_names=names, vocab_words=vocab_words, wv=wv,
                       relations=relations, train=train, dev=labelled(n_dev), test=labelled(n_test))


def make_shaped(shape: str, d: int = 100, Nw: int = 67447, T: int | None = None, seed: int = 0,
                with_eval: bool = False) -> SyntheticKB:
    Ne, R, T_full, n_dev, n_test = SHAPES[shape]
    return make_kb(Ne, R, T_full if T is None else T, Nw, d,
                   n_dev if with_eval else 0, n_test if with_eval else 0, seed=seed)


def make_structured_kb(Ne=600, R=4, T=3000, d=16, clusters=12, n_dev=400, n_test=400, seed=0) -> SyntheticKB:
    """A small KB with learnable structure, for the train -> threshold -> accuracy flow: every entity
    belongs to a latent cluster and is named after its cluster's word plus a private word; relation r
    links cluster c to cluster (a_r*c + b_r) mod clusters.  Positives follow the rule, negatives do not."""
    rng = np.random.default_rng(seed)
    cl = rng.integers(0, clusters, size=Ne)
    vocab_words = [f"c{i}" for i in range(clusters)] + [f"p{i}" for i in range(Ne)] + ["unknown"]
    names = [f"__c{cl[n]}_p{n}_{1 + n % 9}" for n in range(Ne)]
    Nw = len(vocab_words)
    wv = (np.random.default_rng(seed + 1).standard_normal((d, Nw)) * 0.1).astype(np.float32)
    a = rng.integers(1, clusters, size=R) | 1
    b = rng.integers(0, clusters, size=R)
    members = [np.nonzero(cl == c)[0] for c in range(clusters)]

    def sample(n, positive):
        rel = rng.integers(0, R, size=n)
        e1 = rng.integers(0, Ne, size=n)
        tgt = (a[rel] * cl[e1] + b[rel]) % clusters
        if not positive:
            tgt = (tgt + rng.integers(1, clusters, size=n)) % clusters
        e2 = np.array([rng.choice(members[t]) if len(members[t]) else rng.integers(0, Ne) for t in tgt])
        return np.stack([e1, rel, e2], 1).astype(np.int32)

    def labelled(n):
        pos, neg = sample(n // 2, True), sample(n - n // 2, False)
        t = np.concatenate([pos, neg])
        lab = np.r_[np.ones(len(pos), np.int32), -np.ones(len(neg), np.int32)]
        perm = rng.permutation(n)
        return np.concatenate([t, lab[:, None]], 1)[perm].astype(np.int32)

    return SyntheticKB(d=d, entity_names=names, vocab_words=vocab_words, wv=wv,
                       relations=[f"_rel{i}" for i in range(R)], train=sample(T, True),
                       dev=labelled(n_dev), test=labelled(n_test))
