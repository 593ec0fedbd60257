"""Host-side mirror of the reference's DataFactory (/root/reference/src/DataFactory.java) for the
cost+gradient path: entity-name parsing -> entity/word maps, the batch job with seed-exact corrupt
entities, and the loaders of the Socher text format.  Method names follow the reference.

This module is product code (it feeds the C-ABI library); it does not import the oracle.
"""
from __future__ import annotations

import numpy as np

_MULT, _ADD, _MASK = 0x5DEECE66D, 0xB, (1 << 48) - 1


class JavaRandom:
    """java.util.Random (JDK spec): the reference draws corrupt entities with
    ``new Random().nextInt(numOfentities)`` (DataFactory.java:119,128) and corrupt sides with
    ``Math.random()`` (NTN.java:146).  Seeded here so that runs are reproducible and indices are
    bit-exact against a seeded JVM."""

    def __init__(self, seed: int):
        self.s = (seed ^ _MULT) & _MASK

    def _next(self, bits):
        self.s = (self.s * _MULT + _ADD) & _MASK
        v = self.s >> (48 - bits)
        return v - (1 << 32) if v & 0x80000000 else v

    def nextInt(self, bound=None):
        if bound is None:
            return self._next(32)
        if bound <= 0:
            raise ValueError("bound must be positive")
        r = self._next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if ((u - r + m) & 0xFFFFFFFF) < 0x80000000:
                return r
            u = self._next(31)

    def nextDouble(self):
        return ((self._next(26) << 27) + self._next(27)) * (1.0 / (1 << 53))

    def nextInts(self, n: int, bound: int) -> np.ndarray:
        """n consecutive nextInt(bound) draws (same stream as n calls of nextInt).  Uses the C++ host's
        JavaRandom when the library is built (200k draws per batch job), else the Python loop below."""
        out = np.empty(n, np.int32)
        try:
            import ctypes as C
            from . import _capi
            lib = _capi.load()
            st = C.c_uint64(self.s)
            if lib.ntn_host_random_ints(C.byref(st), int(bound), int(n), out.ctypes.data) == 0:
                self.s = int(st.value)
                return out
        except (ImportError, OSError, AttributeError):
            pass
        s = self.s
        m = bound - 1
        pow2 = bound & m == 0
        for i in range(n):
            while True:
                s = (s * _MULT + _ADD) & _MASK
                u = s >> 17
                if pow2:
                    out[i] = (bound * u) >> 31
                    break
                r = u % bound
                if u - r + m < 0x80000000:
                    out[i] = r
                    break
        self.s = s
        return out


def _java_split(s: str, sep: str = "_"):
    parts = s.split(sep)
    while len(parts) > 1 and parts[-1] == "":
        parts.pop()
    if len(parts) == 1 and parts[0] == "" and s != "":
        return []
    return parts


class DataFactory:
    """DataFactory.getInstance(batch_size, corrupt_size, embedding_size, reduce_RelationSize, german)
    (DataFactory.java:70-84).  ``seed`` replaces the reference's unseeded ``new Random()``."""

    def __init__(self, batch_size, corrupt_size, embedding_size, reduce_RelationSize=False, german=False, seed=42,
                 shuffle_seed=None):
        self.batch_size, self.corrupt_size, self.embeddings_size = batch_size, corrupt_size, embedding_size
        self.reduced_RelationSize, self.german = reduce_RelationSize, german
        self.rand = JavaRandom(seed)
        # Collections.shuffle draws from its own Random in the JDK, so the shuffle has its own seeded stream here
        self.shuffle_rand = JavaRandom(seed + 1 if shuffle_seed is None else shuffle_seed)
        self.entitiesNumWord: list[str] = []
        self.entitiesWordNum: dict[str, int] = {}
        self.relationsNumWord: list[str] = []
        self.relationsWordNum: dict[str, int] = {}
        self.vocabNumWord: list[str] = []
        self.vocabWordNum: dict[str, int] = {}
        self.wordVectorMaxtrixLoaded = None          # d x Nw float32
        self.trainingTripples = np.zeros((0, 3), np.int32)
        self.devTripples = np.zeros((0, 4), np.int32)
        self.testTripples = np.zeros((0, 4), np.int32)
        self.batchjob = None                          # (e1, rel, e2, e3) int32 columns
        self._maps = None

    # ---- in-memory population (synthetic data) -------------------------------------------
    @classmethod
    def from_kb(cls, kb, batch_size, corrupt_size, seed=42):
        f = cls(batch_size, corrupt_size, kb.d, seed=seed)
        f.setEntities(kb.entity_names)
        f.setRelations(kb.relations)
        f.setWordVectors(kb.vocab_words, kb.wv)
        f.trainingTripples = np.ascontiguousarray(kb.train, np.int32)
        f.devTripples = np.ascontiguousarray(kb.dev, np.int32)
        f.testTripples = np.ascontiguousarray(kb.test, np.int32)
        return f

    def setEntities(self, names):
        self.entitiesNumWord = list(names)
        self.entitiesWordNum = {n: i for i, n in enumerate(self.entitiesNumWord)}
        self._maps = None

    def setRelations(self, names):
        self.relationsNumWord = list(names)
        self.relationsWordNum = {n: i for i, n in enumerate(self.relationsNumWord)}

    def setWordVectors(self, words, wv):
        self.vocabNumWord = list(words)
        self.vocabWordNum = {w: i for i, w in enumerate(self.vocabNumWord)}
        self.wordVectorMaxtrixLoaded = np.ascontiguousarray(wv, np.float32)
        assert self.wordVectorMaxtrixLoaded.shape == (self.embeddings_size, len(self.vocabNumWord))
        self._maps = None

    # ---- loaders of the Socher text files (DataFactory.java:153-325) -----------------------
    def loadEntitiesFromSocherFile(self, path):
        with open(path) as f:
            self.setEntities([ln.rstrip("\n") for ln in f if ln.rstrip("\n") != ""])

    def loadRelationsFromSocherFile(self, path):
        with open(path) as f:
            names = [ln.rstrip("\n") for ln in f if ln.rstrip("\n") != ""]
        if self.reduced_RelationSize:
            names = names[:2]                                   # reduceRelationToSize = 2 (:62,206)
        self.setRelations(names)

    def _load_triples(self, path, labelled):
        rows = []
        with open(path) as f:
            for ln in f:
                t = ln.split()                                  # line.split("\\s") (:227-229,262-265)
                if not t:
                    continue
                if t[1] not in self.relationsWordNum:
                    if self.reduced_RelationSize:
                        continue
                    raise KeyError(t[1])
                row = [self.entitiesWordNum[t[0]], self.relationsWordNum[t[1]], self.entitiesWordNum[t[2]]]
                if labelled:
                    row.append(int(t[3]))
                rows.append(row)
        return np.asarray(rows, np.int32).reshape(-1, 4 if labelled else 3)

    def loadTrainingDataTripplesE1rE2(self, path):
        self.trainingTripples = self._load_triples(path, False)

    def loadDevDataTripplesE1rE2Label(self, path):
        self.devTripples = self._load_triples(path, True)

    def loadTestDataTripplesE1rE2Label(self, path):
        self.testTripples = self._load_triples(path, True)

    def loadWordVectorsFromMatFile(self, path, optimizedLoad=False):
        """initEmbed.mat: cell ``words`` + matrix ``We`` read as Nw x d (DataFactory.java:327-391,
        SURVEY.md App. B15)."""
        from scipy.io import loadmat
        m = loadmat(path)
        We = np.asarray(m["We"], np.float64)
        words = [str(np.ravel(w)[0]) for w in np.ravel(m["words"])]
        if We.shape[0] != len(words):
            We = We.T
        self.setWordVectors(words, We.T.astype(np.float32))

    # ---- sizes ----------------------------------------------------------------------------
    def getNumOfentities(self):
        return len(self.entitiesNumWord)

    def getNumOfRelations(self):
        return len(self.relationsNumWord)

    def getNumOfWords(self):
        return len(self.vocabNumWord)

    def getWordVectorMaxtrixLoaded(self):
        return self.wordVectorMaxtrixLoaded

    def getAllTrainingTripples(self):
        return self.trainingTripples

    def getDevTripples(self):
        return self.devTripples

    def getTestTripples(self):
        return self.testTripples

    # ---- entity names -> words (DataFactory.java:397-434,583-595,646-684) -------------------
    @staticmethod
    def _clear_name(raw):
        last = raw.rfind("_")
        return raw[2:last] if last >= 2 else raw[2:]

    def _word_num(self, w):
        i = self.vocabWordNum.get(w)
        return self.vocabWordNum["unknown"] if i is None else i

    def entityLength(self, i):
        return len(_java_split(self.entitiesNumWord[i])) - 3

    def getWordIndexes(self, i):
        n = self.entityLength(i)
        out = [0] * (n if n != 0 else 1)
        name = self._clear_name(self.entitiesNumWord[i])
        words = _java_split(name) if "_" in name else [name]
        for j, w in enumerate(words):
            out[j] = self._word_num(w)
        return out

    def entityWordMaps(self):
        """(fwd_off, fwd_idx, bwd_off, bwd_idx, len_bwd): forward lists are what
        createVectorsForEachEntityByWordVectors averages (:405-434); backward lists are
        getWordIndexes / entityLength (what NTN.java:414-426 uses)."""
        if self._maps is None:
            fo, fi, bo, bi, lb = [0], [], [0], [], []
            for i, raw in enumerate(self.entitiesNumWord):
                name = self._clear_name(raw)
                words = _java_split(name) if "_" in name else [name]
                fi.extend(self._word_num(w) for w in words)
                fo.append(len(fi))
                bi.extend(self.getWordIndexes(i))
                bo.append(len(bi))
                lb.append(self.entityLength(i))
            a = lambda x: np.asarray(x, np.int32)
            self._maps = (a(fo), a(fi), a(bo), a(bi), a(lb))
        return self._maps

    # ---- batch job (DataFactory.java:116-136) -----------------------------------------------
    def shuffleTrainingTripples(self):
        """Collections.shuffle(trainingTripples) -- the line the reference has commented out (DataFactory.java:118,
        "TODO wieder zulassen") -- in place, JDK algorithm: for i = size..2: swap(i-1, rnd.nextInt(i))."""
        n = len(self.trainingTripples)
        draws = [self.shuffle_rand.nextInt(i) for i in range(n, 1, -1)]
        perm = np.arange(n)
        for i, j in zip(range(n, 1, -1), draws):
            perm[i - 1], perm[j] = perm[j], perm[i - 1]
        self.trainingTripples = np.ascontiguousarray(self.trainingTripples[perm])
        return perm

    def generateNewTrainingBatchJob(self, first=0, shuffle=False):
        """h-major: for h < corrupt_size, for i < batch_size: e3 = rand.nextInt(numOfentities);
        always training triples [first, first+batch_size) -- first = 0 in the reference (the
        shuffle is commented out, :118); a non-zero ``first`` is how a data-parallel rank takes
        its shard of a larger batch.  ``shuffle=True`` re-enables the commented-out shuffle."""
        if shuffle:
            self.shuffleTrainingTripples()
        tr = self.trainingTripples[first:first + self.batch_size]
        n = len(tr)
        e3 = self.rand.nextInts(n * self.corrupt_size, self.getNumOfentities())
        rep = lambda c: np.ascontiguousarray(np.tile(tr[:, c], self.corrupt_size), np.int32)
        self.batchjob = (rep(0), rep(1), rep(2), e3)
        return self.batchjob

    def getBatchJobTripplesOfRelation(self, r):
        e1, rel, e2, e3 = self.batchjob
        sel = rel == r
        return e1[sel], rel[sel], e2[sel], e3[sel]

    def getTripplesOfRelation(self, r, tripples):
        return tripples[tripples[:, 1] == r]
