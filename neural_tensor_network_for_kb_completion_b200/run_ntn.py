"""Run_NTN: the reference's driver flow (/root/reference/src/Run_NTN.java:38-150) over the B200 path.

    python -m neural_tensor_network_for_kb_completion_b200.run_ntn <data_path> <theta_load_path> <theta_save_path> [numIterations]

Same hyper-parameters as the reference's hard-coded locals (Run_NTN.java:72-83), same loop:
for every outer iteration build a new batch job, run L-BFGS with maxIters = batch_iterations from the
current theta (fresh minimiser each time, :119-124), checkpoint theta as comma-separated text when
i == 5 or i % 10 == 0 (:133-135), then thresholds on dev and accuracy on test (:145-149).
"""
from __future__ import annotations

import os
import sys
import time
from dataclasses import dataclass

import numpy as np

from .data_factory import DataFactory
from .lbfgs import LBFGSMinimizer, Opts
from .ntn import NTN


@dataclass
class Config:                     # Run_NTN.java:72-83
    batchSize: int = 20000
    numWVdimensions: int = 100
    numIterations: int = 500
    batch_iterations: int = 5
    sliceSize: int = 3
    corrupt_size: int = 10
    activation_function: str = "tanh"
    lamda: float = 0.0001
    optimizedLoad: bool = False
    reducedNumOfRelations: bool = False
    wv_source: str = "frozen"     # the reference's behaviour
    seed: int = 42


def write_theta_txt(theta, path):
    """Nd4j.writeTxt(theta, path, ","): one comma-separated line of P floats (Run_NTN.java:134,139).  Every value is
    written with enough digits to round-trip its FP32 bits."""
    with open(path, "w") as f:
        f.write(",".join(repr(float(np.float32(v))) for v in theta))


def read_theta_txt(path):
    """The resume the reference only has commented out (Run_NTN.java:108)."""
    with open(path) as f:
        return np.array([float(x) for x in f.read().strip().split(",")], np.float64)


_BIN_MAGIC = b"NTNTHETA"


def write_theta_bin(theta, path):
    """Binary fast path beside the text checkpoint: "NTNTHETA" | uint32 version = 1 | uint32 dtype = 32 | int64 P |
    P little-endian float32 (same layout as csrc/host/ntn_host.hpp writeThetaBin)."""
    v = np.asarray(theta, np.float32).astype("<f4")
    with open(path, "wb") as f:
        f.write(_BIN_MAGIC + np.array([1, 32], "<u4").tobytes() + np.array([v.size], "<i8").tobytes())
        v.tofile(f)


def read_theta(path):
    """Either checkpoint format (the binary one is recognised by its magic) as float64 values of FP32 numbers."""
    with open(path, "rb") as f:
        head = f.read(8)
        if head == _BIN_MAGIC:
            ver, dt = np.frombuffer(f.read(8), "<u4")
            n = int(np.frombuffer(f.read(8), "<i8")[0])
            if ver != 1 or dt != 32:
                raise ValueError(f"unsupported theta checkpoint header in {path}")
            v = np.fromfile(f, "<f4", n)
            if v.size != n:
                raise ValueError(f"truncated theta checkpoint {path}")
            return v.astype(np.float64)
    return read_theta_txt(path).astype(np.float32).astype(np.float64)


def save_checkpoint(theta, path_txt):
    write_theta_txt(theta, path_txt)
    write_theta_bin(theta, os.path.splitext(path_txt)[0] + ".bin")


def load_data(data_path, cfg: Config) -> DataFactory:
    tbj = DataFactory(cfg.batchSize, cfg.corrupt_size, cfg.numWVdimensions, cfg.reducedNumOfRelations, False, seed=cfg.seed)
    tbj.loadEntitiesFromSocherFile(os.path.join(data_path, "entities.txt"))      # Run_NTN.java:91-95
    tbj.loadRelationsFromSocherFile(os.path.join(data_path, "relations.txt"))
    tbj.loadTrainingDataTripplesE1rE2(os.path.join(data_path, "train.txt"))
    tbj.loadDevDataTripplesE1rE2Label(os.path.join(data_path, "dev.txt"))
    tbj.loadTestDataTripplesE1rE2Label(os.path.join(data_path, "test.txt"))
    tbj.loadWordVectorsFromMatFile(os.path.join(data_path, "initEmbed.mat"), cfg.optimizedLoad)
    return tbj


def train_and_evaluate(tbj: DataFactory, cfg: Config, theta=None, theta_save_path=None, log=print, minimizer=None):
    """Run_NTN.java:103-149.  ``minimizer(t, theta, opts) -> Result`` defaults to the device L-BFGS."""
    t = NTN(cfg.numWVdimensions, tbj.getNumOfentities(), tbj.getNumOfRelations(), tbj.getNumOfWords(), cfg.batchSize,
            cfg.sliceSize, cfg.activation_function, tbj, cfg.lamda, wv_source=cfg.wv_source)
    t.connectDatafactory(tbj)
    theta = np.asarray(t.getTheta_inital(), np.float64) if theta is None else np.asarray(theta, np.float64)
    minimize = minimizer or (lambda fn, th, o: LBFGSMinimizer().minimize(fn, th, o))
    history = []
    for i in range(cfg.numIterations):                                           # :113
        tbj.generateNewTrainingBatchJob()                                        # :115
        opts = Opts(maxIters=cfg.batch_iterations)                               # :119-121
        t0 = time.perf_counter()
        res = minimize(t, theta, opts)                                           # :122
        log(f"res: {res.didConverge}| {res.minObjVal}")                          # :123
        theta = res.minArg                                                       # :124
        history.append((res.firstObjVal, res.minObjVal, res.evals, time.perf_counter() - t0))
        log(f"Paramters for batchjob optimized, iteration: {i} completed")       # :131
        if theta_save_path and (i == 5 or i % 10 == 0):                          # :133-135
            save_checkpoint(theta, os.path.join(theta_save_path, f"theta_opt_iteration_{i}.txt"))
    if theta_save_path:                                                          # :139
        save_checkpoint(theta, os.path.join(theta_save_path, f"theta_opt{time.localtime().tm_mday}.txt"))
        log("Model saved!")
    best_theresholds = t.computeBestThresholds(theta, tbj.getDevTripples())      # :145
    log(f"Best theresholds: {best_theresholds}")
    predictions, accuracy = t.getPrediction(theta, tbj.getTestTripples(), best_theresholds)   # :149
    _, dev_accuracy = t.getPrediction(theta, tbj.getDevTripples(), best_theresholds)
    t.close()
    return {"theta": theta, "thresholds": best_theresholds, "predictions": predictions, "accuracy": accuracy,
            "dev_accuracy": dev_accuracy, "history": history}


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 3:
        print(__doc__)
        return 2
    data_path, theta_load_path, theta_save_path = argv[:3]                       # Run_NTN.java:61-69
    cfg = Config()
    if len(argv) > 3:
        cfg.numIterations = int(argv[3])
    print(f"NTN: batchSize: {cfg.batchSize} | SliceSize: {cfg.sliceSize} | numIterations:{cfg.numIterations} | "
          f"corrupt_size: {cfg.corrupt_size}| activation func: {cfg.activation_function}")
    tbj = load_data(data_path, cfg)
    theta = read_theta(theta_load_path) if os.path.isfile(theta_load_path) else None   # resume (Run_NTN.java:64,108)
    train_and_evaluate(tbj, cfg, theta, theta_save_path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
