"""Host-side mirror of the reference's NTN function object (/root/reference/src/NTN.java).

Same constructor arguments, same method names (computeAt / getDimension / valueAt /
derivativeAt / domainDimension / computeBestThresholds / getPrediction / getTheta_inital), same
flattened-theta contract -- but every number is computed by the CUDA library behind
include/ntn_b200.h.  No CPU fallback: constructing an NTN without the library or without a GPU
raises.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from .data_factory import DataFactory, JavaRandom


class NTN:
    def __init__(self, embeddingSize, numberOfEntities, numberOfRelations, numberOfWords, batchSize,
                 sliceSize, activation_function, tbj: DataFactory, lamda,
                 wv_source="frozen", gemm_path="auto", device=0, reg_scale=1.0, side_seed=7, init_seed=3):
        """NTN.java:42-84.  Extra keyword arguments are the knobs the reference leaves implicit:
        wv_source  "frozen" = the reference's behaviour (entity vectors from the load-time word
                   vectors), "theta" = from theta's word block;
        side_seed  seed of the stream standing in for Math.random() at NTN.java:146."""
        self.embeddingSize, self.numberOfEntities = embeddingSize, numberOfEntities
        self.numberOfRelations, self.numOfWords = numberOfRelations, numberOfWords
        self.batchSize, self.sliceSize, self.lamda = batchSize, sliceSize, float(lamda)
        self.activationFunc = activation_function
        self.update = 0                                            # NTN.java:39,66
        self.tbj = tbj
        self.device = device
        self.side_rand = JavaRandom(side_seed)
        self.handle = _capi.Handle(embeddingSize, sliceSize, numberOfRelations, numberOfEntities, numberOfWords,
                                   batchSize, self.lamda, activation=activation_function,
                                   wv_source=wv_source, gemm_path=gemm_path, device=device, reg_scale=reg_scale)
        self.handle.set_entity_words(*tbj.entityWordMaps())
        wv = tbj.getWordVectorMaxtrixLoaded()
        if wv_source == "frozen":
            self.handle.set_frozen_wv(wv)
        # NTN.java:67-82: W = rand * r, r = 1/sqrt(2d) (App. B11); V = 0.4, b = 0.01, U = 0.8
        d, k, R = embeddingSize, sliceSize, numberOfRelations
        rng = np.random.default_rng(init_seed)
        r = 1.0 / np.sqrt(2.0 * d)
        self.theta_inital = np.concatenate([
            (rng.random(R * k * d * d) * r), np.full(R * 2 * d * k, 0.4), np.full(R * k, 0.01),
            np.full(R * k, 0.8), np.asarray(wv, np.float64).ravel()]).astype(np.float32)
        self.dimension_for_minimizer = self.theta_inital.size
        assert self.dimension_for_minimizer == self.handle.P
        self._batch_id = None
        self.last_corrupt_side = None
        self._present_for, self._present = None, None
        self._train_id = None
        self.global_relations = None     # data parallel: relation ids of the GLOBAL batch job (see drawCorruptSides)

    # ---- DataFactory hand-off ----------------------------------------------------------------
    def connectDatafactory(self, tbj):
        self.tbj = tbj

    def getDatafactory(self):
        return self.tbj

    def getTheta_inital(self):
        return self.theta_inital

    def _sync_batch(self):
        bj = self.tbj.batchjob
        if bj is None:
            raise RuntimeError("generateNewTrainingBatchJob() has not been called")
        if self._batch_id is not bj:                               # DataFactory mutated its batch job
            self.handle.set_batch(*bj)
            self._batch_id = bj

    def generateBatchJobOnDevice(self, seed, job, first=0, filter_true=False, fetch=False):
        """Sampler variant: the batch job over training triples [first, first+batch) is generated and
        indexed in HBM -- corrupt entities from the counter-based device sampler, optionally redrawn while they
        reproduce a known training triple.  Not the parity mode (that is DataFactory.generateNewTrainingBatchJob with its
        seeded java.util.Random).  ``fetch`` downloads the columns into tbj.batchjob (tests, inspection)."""
        tr = self.tbj.getAllTrainingTripples()
        if self._train_id is not tr:
            self.handle.set_training_triples(tr)
            self._train_id = tr
        B, C = self.tbj.batch_size, self.tbj.corrupt_size
        B = min(B, len(tr) - first)
        self.handle.generate_batch_dev(first, B, C, seed, job, filter_true)
        if fetch:
            self.tbj.batchjob = self.handle.get_batch_columns(B * C)
        else:                                                      # only the relation column is needed on the host
            self.tbj.batchjob = (None, np.ascontiguousarray(tr[first:first + B, 1]), None, None)
        self._batch_id = self.tbj.batchjob                         # already on the device: nothing to sync
        return self.tbj.batchjob

    def drawCorruptSides(self):
        """One Math.random() > 0.5 per non-empty relation, increasing r (NTN.java:120,146).  A data-parallel rank holds
        a shard that may miss a rare relation: ``global_relations`` (the global batch job's relation column) keeps the
        number of draws -- hence the seeded stream -- identical on every rank."""
        rel = self.tbj.batchjob[1] if self.global_relations is None else self.global_relations
        if self._present_for is not rel:                           # one pass per batch job, not per evaluation
            self._present = np.bincount(np.asarray(rel), minlength=self.numberOfRelations)[:self.numberOfRelations] > 0
            self._present_for = rel
        side = np.zeros(self.numberOfRelations, np.uint8)
        for r in np.nonzero(self._present)[0]:
            side[r] = 1 if self.side_rand.nextDouble() > 0.5 else 0
        return side

    def nativeSidesState(self):
        """State array for the library's native corrupt-side drawer (ntn_host_sides_draw): the same seeded stream as
        drawCorruptSides, advanced in C while the device L-BFGS runs; ``adoptSidesState`` takes the stream back."""
        rel = self.tbj.batchjob[1] if self.global_relations is None else self.global_relations
        if self._present_for is not rel:
            self._present = np.bincount(np.asarray(rel), minlength=self.numberOfRelations)[:self.numberOfRelations] > 0
            self._present_for = rel
        st = np.zeros(2 + self.numberOfRelations, np.uint64)
        st[0] = self.side_rand.s
        st[1] = self.numberOfRelations
        st[2:] = self._present.astype(np.uint64)
        return st

    def adoptSidesState(self, st):
        self.side_rand.s = int(st[0])

    # ---- IDifferentiableFn (NTN.java:92-448) ----------------------------------------------------
    def computeAt(self, _theta, corrupt_side=None):
        """IPair<Double,double[]> computeAt(double[] theta) -> (cost, grad)."""
        self._sync_batch()
        side = self.drawCorruptSides() if corrupt_side is None else corrupt_side
        self.last_corrupt_side = side
        cost, grad, nact = self.handle.eval(_theta, side)
        self.update += nact                                        # NTN.java:222
        return cost, grad

    def getDimension(self):
        return self.dimension_for_minimizer

    # ---- DiffFunction (NTN.java:770-785) --------------------------------------------------------
    def domainDimension(self):
        return self.dimension_for_minimizer

    def valueAt(self, theta):
        return self.computeAt(theta)[0]

    def derivativeAt(self, theta):
        return self.computeAt(theta)[1]

    # ---- evaluation (NTN.java:587-736) ------------------------------------------------------------
    def scoreTripples(self, _theta, tripples, mode="reference_eval"):
        t = np.asarray(tripples)
        return self.handle.score(_theta, t[:, 0], t[:, 1], t[:, 2], mode)

    def computeBestThresholds(self, _theta, _devTrippels):
        """NTN.java:587-669, quirks kept (App. B7): thresholds start at -1, ``score_temp += 0.01``
        advances inside the relation loop, strict improvement only."""
        dev = np.asarray(_devTrippels)
        scores = self.scoreTripples(_theta, dev)
        R = self.numberOfRelations
        best_thr, best_acc = np.full(R, -1.0), np.zeros(R)
        per_rel = [np.nonzero(dev[:, 1] == i)[0] for i in range(R)]
        t, smax = float(scores.min()), float(scores.max())
        while t <= smax:
            for i in range(R):
                idx = per_rel[i]
                if idx.size:
                    pt = scores[idx] <= t
                    acc = (np.sum(pt & (dev[idx, 3] == 1)) + np.sum(~pt & (dev[idx, 3] == -1))) / idx.size
                    if acc > best_acc[i]:
                        best_acc[i], best_thr[i] = acc, t
                t = t + 0.01
        return best_thr

    def getPrediction(self, _theta, _testTripples, _bestThresholds):
        """NTN.java:672-712 -> (predictions, accuracy)."""
        test = np.asarray(_testTripples)
        scores = self.scoreTripples(_theta, test)
        pred = np.where(scores <= np.asarray(_bestThresholds)[test[:, 1]], 1, -1)
        acc = float(np.mean(pred == test[:, 3]))
        print("Accuracy of predictions: %s" % acc)               # NTN.java:709
        return pred, acc

    def close(self):
        self.handle.close()
