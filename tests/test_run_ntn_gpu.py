"""End-to-end Run_NTN flow on the GPU (train -> thresholds -> accuracy) against the same flow driven by the
CPU oracle (same seeds, same L-BFGS definition, same threshold sweep): final dev/test accuracy within the
north-star band.  Also the loaders / checkpoint formats (next rows N2, N3) on files written in the Socher
layout."""
import os

import numpy as np
import pytest

from neural_tensor_network_for_kb_completion_b200 import run_ntn, synth
from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory, JavaRandom
from oracle import data_ref, lbfgs_ref, ntn_ref
from oracle.java_random import JavaRandom as OracleRandom

pytestmark = pytest.mark.gpu


def oracle_flow(kb, cfg):
    """Run_NTN.java:103-149 with every number computed by the oracle (float64)."""
    vocab = {w: i for i, w in enumerate(kb.vocab_words)}
    fo, fi, bo, bi, lb = data_ref.build_entity_word_maps(kb.entity_names, vocab)
    p = ntn_ref.NTNProblem(d=kb.d, k=cfg.sliceSize, R=kb.R, Ne=kb.Ne, Nw=kb.Nw, lam=cfg.lamda, batch_size=cfg.batchSize,
                           fwd_off=fo, fwd_idx=fi, bwd_off=bo, bwd_idx=bi, len_bwd=lb)
    e3_rand, side_rand = OracleRandom(cfg.seed), OracleRandom(7)
    # same theta0 as the NTN mirror (init_seed=3)
    rng = np.random.default_rng(3)
    r = 1.0 / np.sqrt(2.0 * kb.d)
    theta = np.concatenate([rng.random(kb.R * cfg.sliceSize * kb.d * kb.d) * r, np.full(kb.R * 2 * kb.d * cfg.sliceSize, 0.4),
                            np.full(kb.R * cfg.sliceSize, 0.01), np.full(kb.R * cfg.sliceSize, 0.8),
                            kb.wv.astype(np.float64).ravel()]).astype(np.float32).astype(np.float64)
    wvf = kb.wv if cfg.wv_source == "frozen" else None
    for _ in range(cfg.numIterations):
        batch = data_ref.generate_batch_job(kb.train[:, 0], kb.train[:, 1], kb.train[:, 2], cfg.batchSize, cfg.corrupt_size,
                                            kb.Ne, e3_rand)

        def fn(th):
            side = data_ref.draw_corrupt_sides(batch[1], kb.R, side_rand)
            return ntn_ref.compute_at(th, p, *batch, side, wv_frozen=wvf)
        theta = lbfgs_ref.minimize(fn, theta, cfg.batch_iterations)[0]
    s_dev = ntn_ref.score_triples_reference_eval(theta, p, kb.dev[:, 0], kb.dev[:, 1], kb.dev[:, 2], wv_frozen=wvf)
    thr, _ = ntn_ref.compute_best_thresholds(s_dev, kb.dev[:, 1], kb.dev[:, 3], kb.R)
    s_test = ntn_ref.score_triples_reference_eval(theta, p, kb.test[:, 0], kb.test[:, 1], kb.test[:, 2], wv_frozen=wvf)
    return (ntn_ref.get_prediction(s_dev, kb.dev[:, 1], kb.dev[:, 3], thr)[1],
            ntn_ref.get_prediction(s_test, kb.test[:, 1], kb.test[:, 3], thr)[1], theta)


@pytest.mark.parametrize("wv_source", ["theta", "frozen"])
def test_train_threshold_accuracy_flow_matches_oracle_flow(wv_source):
    kb = synth.make_structured_kb(Ne=400, R=3, T=1500, d=12, clusters=8, n_dev=5000, n_test=5000, seed=1)
    cfg = run_ntn.Config(batchSize=1000, numWVdimensions=12, numIterations=6, batch_iterations=5, sliceSize=2,
                         corrupt_size=4, lamda=1e-4, wv_source=wv_source, seed=42)
    tbj = DataFactory.from_kb(kb, cfg.batchSize, cfg.corrupt_size, seed=cfg.seed)
    out = run_ntn.train_and_evaluate(tbj, cfg, log=lambda *_: None)
    dev_ref, test_ref, theta_ref = oracle_flow(kb, cfg)
    first, last = out["history"][0][0], out["history"][-1][1]
    assert last < first                                            # training reduces the objective
    if wv_source == "theta":
        assert out["accuracy"] > 0.6                               # and the structure is learnt
    # north star: final dev/test accuracy within 0.2 points of the reference flow (5,000 labelled triples each: one
    # flipped prediction is 0.02 point; the FP32 trajectory against the oracle's FP64 one may flip a few)
    assert abs(out["dev_accuracy"] - dev_ref) <= 0.002, (out["dev_accuracy"], dev_ref)
    assert abs(out["accuracy"] - test_ref) <= 0.002, (out["accuracy"], test_ref)


def test_wordnet_sized_training_then_accuracy_by_gpu_and_by_oracle():
    """BASELINE.json configs[1] at its real size (38,696 entities, 11 relations, 112,581 training triples, d = 100, k = 3,
    batch 20,000, corrupt 10): the whole Run_NTN flow on the GPU -- 40 batch jobs x L-BFGS(5) -- and then the threshold
    sweep and the dev / test accuracy of the SAME trained theta computed twice: by the GPU scoring path and by the
    float64 oracle (NTN.java:587-736).  The north star's 0.2-point band applies to that pair; the trajectory itself is
    compared with the oracle's on the small KB above (a full-size oracle trajectory is days of CPU)."""
    kb = synth.make_structured_kb(Ne=38696, R=11, T=112581, d=100, clusters=24, n_dev=5000, n_test=10000, seed=3)
    cfg = run_ntn.Config(batchSize=20000, numWVdimensions=100, numIterations=40, batch_iterations=5, sliceSize=3,
                         corrupt_size=10, lamda=1e-4, wv_source="theta", seed=42)
    tbj = DataFactory.from_kb(kb, cfg.batchSize, cfg.corrupt_size, seed=cfg.seed)
    out = run_ntn.train_and_evaluate(tbj, cfg, log=lambda *_: None)
    assert out["history"][-1][1] < out["history"][0][0]
    assert out["accuracy"] > 0.6                                   # the cluster structure is learnt at this size too
    theta = np.asarray(out["theta"], np.float64)
    vocab = {w: i for i, w in enumerate(kb.vocab_words)}
    fo, fi, bo, bi, lb = data_ref.build_entity_word_maps(kb.entity_names, vocab)
    p = ntn_ref.NTNProblem(d=kb.d, k=cfg.sliceSize, R=kb.R, Ne=kb.Ne, Nw=kb.Nw, lam=cfg.lamda, batch_size=cfg.batchSize,
                           fwd_off=fo, fwd_idx=fi, bwd_off=bo, bwd_idx=bi, len_bwd=lb)
    s_dev = ntn_ref.score_triples_reference_eval(theta, p, kb.dev[:, 0], kb.dev[:, 1], kb.dev[:, 2])
    thr, _ = ntn_ref.compute_best_thresholds(s_dev, kb.dev[:, 1], kb.dev[:, 3], kb.R)
    s_test = ntn_ref.score_triples_reference_eval(theta, p, kb.test[:, 0], kb.test[:, 1], kb.test[:, 2])
    dev_ref = ntn_ref.get_prediction(s_dev, kb.dev[:, 1], kb.dev[:, 3], thr)[1]
    test_ref = ntn_ref.get_prediction(s_test, kb.test[:, 1], kb.test[:, 3], thr)[1]
    assert abs(out["dev_accuracy"] - dev_ref) <= 0.002, (out["dev_accuracy"], dev_ref)
    assert abs(out["accuracy"] - test_ref) <= 0.002, (out["accuracy"], test_ref)
    print(f"configs[1] size: dev accuracy gpu {out['dev_accuracy']:.4f} / oracle {dev_ref:.4f}, "
          f"test accuracy gpu {out['accuracy']:.4f} / oracle {test_ref:.4f}")


def test_socher_files_mat_vectors_and_theta_checkpoints(tmp_path):
    from scipy.io import savemat
    kb = synth.make_structured_kb(Ne=120, R=3, T=300, d=8, clusters=6, n_dev=40, n_test=40, seed=2)
    d = str(tmp_path)
    open(os.path.join(d, "entities.txt"), "w").write("\n".join(kb.entity_names) + "\n")
    open(os.path.join(d, "relations.txt"), "w").write("\n".join(kb.relations) + "\n")
    name = lambda t: f"{kb.entity_names[t[0]]}\t{kb.relations[t[1]]}\t{kb.entity_names[t[2]]}"
    open(os.path.join(d, "train.txt"), "w").write("\n".join(name(t) for t in kb.train) + "\n")
    for fn, arr in (("dev.txt", kb.dev), ("test.txt", kb.test)):
        open(os.path.join(d, fn), "w").write("\n".join(name(t) + f"\t{t[3]}" for t in arr) + "\n")
    words = np.empty((1, kb.Nw), dtype=object)
    for i, w in enumerate(kb.vocab_words):
        words[0, i] = np.array([w])
    savemat(os.path.join(d, "initEmbed.mat"), {"words": words, "We": kb.wv.T.astype(np.float64)})   # Nw x d (App. B15)
    cfg = run_ntn.Config(batchSize=100, numWVdimensions=8, numIterations=2, batch_iterations=2, sliceSize=2, corrupt_size=2)
    tbj = run_ntn.load_data(d, cfg)
    ref = DataFactory.from_kb(kb, 100, 2)
    assert tbj.getNumOfentities() == kb.Ne and tbj.getNumOfRelations() == kb.R and tbj.getNumOfWords() == kb.Nw
    assert np.array_equal(tbj.trainingTripples, kb.train) and np.array_equal(tbj.devTripples, kb.dev)
    assert np.allclose(tbj.getWordVectorMaxtrixLoaded(), kb.wv)
    for a, b in zip(tbj.entityWordMaps(), ref.entityWordMaps()):
        assert np.array_equal(a, b)
    out = run_ntn.train_and_evaluate(tbj, cfg, theta_save_path=d, log=lambda *_: None)
    ck = os.path.join(d, "theta_opt_iteration_0.txt")
    assert os.path.exists(ck) and os.path.exists(ck[:-4] + ".bin")
    # checkpoints round-trip bit for bit in FP32, text (Nd4j.writeTxt layout, Run_NTN.java:134,139) and binary
    import time as _time
    final = os.path.join(d, f"theta_opt{_time.localtime().tm_mday}.txt")
    bits = np.asarray(out["theta"], np.float32).view(np.uint32)
    for path in (final, final[:-4] + ".bin"):
        assert np.array_equal(run_ntn.read_theta(path).astype(np.float32).view(np.uint32), bits), path
    assert np.array_equal(run_ntn.read_theta(ck), run_ntn.read_theta(ck[:-4] + ".bin"))
    # resume through theta_load_path (the load the reference has commented out, Run_NTN.java:64,108): with no further
    # training the resumed run reproduces thresholds, predictions and accuracy of the run that wrote the checkpoint
    cfg0 = run_ntn.Config(batchSize=100, numWVdimensions=8, numIterations=0, batch_iterations=2, sliceSize=2, corrupt_size=2)
    for path in (final, final[:-4] + ".bin"):
        again = run_ntn.train_and_evaluate(run_ntn.load_data(d, cfg0), cfg0, theta=run_ntn.read_theta(path), log=lambda *_: None)
        assert again["accuracy"] == out["accuracy"] and again["dev_accuracy"] == out["dev_accuracy"]
        assert np.array_equal(again["thresholds"], out["thresholds"]) and np.array_equal(again["predictions"], out["predictions"])
    # ... and training continues from it: one more outer iteration starts at the checkpointed parameters
    cfg1 = run_ntn.Config(batchSize=100, numWVdimensions=8, numIterations=1, batch_iterations=2, sliceSize=2, corrupt_size=2)
    more = run_ntn.train_and_evaluate(run_ntn.load_data(d, cfg1), cfg1, theta=run_ntn.read_theta(final), log=lambda *_: None)
    assert np.isfinite(more["history"][0][0]) and more["history"][0][1] <= more["history"][0][0]


def test_host_java_random_is_bit_exact_with_oracle():
    a, b = JavaRandom(42), OracleRandom(42)
    assert a.nextInts(2000, 75043).tolist() == [b.next_int(75043) for _ in range(2000)]
    assert [a.nextDouble() for _ in range(50)] == [b.next_double() for _ in range(50)]
    assert a.nextInts(100, 1024).tolist() == [b.next_int(1024) for _ in range(100)]


@pytest.mark.parametrize("sampler", ["reference", "shuffle", "device", "device-filter"])
def test_run_ntn_binary_with_every_sampler(tmp_path, sampler):
    """The compiled host (csrc/host/run_ntn.cpp = Run_NTN.main, Run_NTN.java:38-150) end to end on Socher-format
    files, once per sampler variant: the reference's seeded java.util.Random job, the re-enabled shuffle, and the
    on-device counter-based sampler without / with true-triple filtering."""
    import subprocess
    import neural_tensor_network_for_kb_completion_b200 as pkg
    exe = os.path.join(os.path.dirname(pkg.__file__), "run_ntn")
    if not os.path.exists(exe):
        pytest.skip("run_ntn binary not built")
    kb = synth.make_structured_kb(Ne=200, R=3, T=500, d=100, clusters=8, n_dev=60, n_test=60, seed=4)
    d = str(tmp_path)
    open(os.path.join(d, "entities.txt"), "w").write("\n".join(kb.entity_names) + "\n")
    open(os.path.join(d, "relations.txt"), "w").write("\n".join(kb.relations) + "\n")
    name = lambda t: f"{kb.entity_names[t[0]]}\t{kb.relations[t[1]]}\t{kb.entity_names[t[2]]}"
    open(os.path.join(d, "train.txt"), "w").write("\n".join(name(t) for t in kb.train) + "\n")
    for fn, arr in (("dev.txt", kb.dev), ("test.txt", kb.test)):
        open(os.path.join(d, fn), "w").write("\n".join(name(t) + f"\t{t[3]}" for t in arr) + "\n")
    with open(os.path.join(d, "wordvectors.txt"), "w") as f:
        for i, w in enumerate(kb.vocab_words):
            f.write(w + " " + " ".join(f"{x:.8g}" for x in kb.wv[:, i]) + "\n")
    r = subprocess.run([exe, d, d, d, "2", sampler], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Accuracy of predictions:" in r.stdout and "Model saved!" in r.stdout
    acc = float(r.stdout.split("Accuracy of predictions:")[1].split()[0])
    assert 0.0 <= acc <= 1.0
    costs = [float(l.split("|")[1]) for l in r.stdout.splitlines() if l.startswith("res:")]
    assert len(costs) == 2 and all(np.isfinite(costs))
    ck = os.path.join(d, "theta_opt_iteration_0.txt")
    assert os.path.exists(ck) and os.path.exists(ck[:-4] + ".bin")
    if sampler == "reference":
        # resume (theta_load_path, Run_NTN.java:64,108) from the final checkpoint, text and binary, with no further training:
        # the same theta gives the same accuracy line
        import time as _time
        final = os.path.join(d, f"theta_opt{_time.localtime().tm_mday}.txt")
        for path in (final, final[:-4] + ".bin"):
            d2 = str(tmp_path / ("resume_" + path[-3:]))
            os.makedirs(d2)
            r2 = subprocess.run([exe, d, path, d2, "0", sampler], capture_output=True, text=True, timeout=240)
            assert r2.returncode == 0, r2.stderr[-2000:]
            assert "theta loaded from" in r2.stdout
            assert float(r2.stdout.split("Accuracy of predictions:")[1].split()[0]) == acc
        # a checkpoint of the wrong dimension is refused
        open(os.path.join(d, "bad.txt"), "w").write("1.0,2.0,3.0")
        r3 = subprocess.run([exe, d, os.path.join(d, "bad.txt"), d, "0", sampler], capture_output=True, text=True, timeout=240)
        assert r3.returncode != 0 and "theta_load_path holds 3 values" in r3.stderr
