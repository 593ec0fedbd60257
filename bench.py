#!/usr/bin/env python
"""bench.py -- NTN cost+gradient evaluations on Freebase-shaped synthetic data (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c1|c3|c4|c5] [--scaling weak|strong]

A "step" is one cost+gradient evaluation (NTN.computeAt, /root/reference/src/NTN.java:92-443)
of one batch job: 20,000 training triples x 10 corrupt copies per GPU (Run_NTN.java:72-77) on
the Freebase shape (75,043 entities, 13 relations, d=100, k=3; BASELINE.json configs[2]).

  value        triples/s, whole job, theta and batch already resident in HBM (ntn_eval_dev)
  e2e          the same through the reference-facing call with HOST buffers: ntn_eval(double* theta, ..., double* grad)
               at N=1; at N>1 the caller lives on rank 0 (host theta -> rank 0 -> NVLink broadcast -> every rank
               evaluates its shard -> gradient exchange -> rank 0 downloads the gradient)
  roofline     the slowest kernel group of the evaluation's main chain, timed with CUDA events on its stream
  cpu_baseline the oracle's float32 restatement of the reference formulation on the host cores
               (N=1, rank 0): one full evaluation
  parity       this run's ntn_eval result against the committed float64-oracle golden vector of the same inputs
               (tests/golden/c3_fullsize.npz), 1e-6 + 1e-4|ref|, pass/fail

N > 1 (torchrun): --scaling weak (default): every rank evaluates its own 20,000 x 10 shard of a G*20,000 triple batch;
--scaling strong: ONE batch job of the config's size (c4: all 316,232 x 10, BASELINE.json configs[3]) sharded over the
ranks.  Either way the gradient is exchanged so that every rank holds the full-batch gradient.
`--impl reference` times the oracle port alone (rank 0 only): full evaluations of the same workload when they fit the
time budget, otherwise the largest prefix of the batch job that does.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE, D, K_SLICES, NW, BATCH, CORRUPT, LAMBDA = "freebase", 100, 3, 67447, 20000, 10, 1e-4
CONFIG_NAME = "BASELINE.json configs[2]"
# headline = c3.  c1: Wordnet-shaped batch (configs[1]); c4: Freebase-shaped FULL batch, all 316,232 triples x 10 (configs[3])
CONFIGS = {"c1": ("wordnet", 20000, "BASELINE.json configs[1]"), "c3": ("freebase", 20000, "BASELINE.json configs[2]"),
           "c4": ("freebase", 316232, "BASELINE.json configs[3]"),
           # configs[4]'s model shape (1M entities, 64 relations, d=256, k=8, 500k words) with a 200,000 x 10 batch job
           "c5": ("scaled", 200000, "BASELINE.json configs[4] shape, 200,000-triple batch job")}
SCALED = dict(Ne=1000000, R=64, d=256, k=8, Nw=500000)
THETA_SEED_DEFAULT = 11
GOLDEN_C3 = os.path.join(ROOT, "tests", "golden", "c3_fullsize.npz")
EVALS_PER_JOB = 22        # cost+gradient evaluations of one L-BFGS(5) call, measured (profiles/r1_train_loop_wordnet.txt: 21.8)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        return float(m["hbm_gbs"]), float(m["bf16_tflops_sustained"]), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


def measure_tf32_tflops(dev, n=8192, iters=8):
    """Dense kind::tf32 tensor throughput of this GPU: cuBLAS FP32 GEMM with TF32 math allowed (tcgen05 kind::tf32 kernels),
    best of `iters`, CUDA events.  The tensor denominator of the 3xTF32 contractions is this / 3."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "25"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.rows.append(l) for l in self.p.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def mark(self):
        """Called when the timed region starts: only samples taken from here on are reported (the sampler itself is
        started before the warm-up so that nvidia-smi's start-up does not eat a short timed region)."""
        self.first = len(self.rows)

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        rows = self.rows[getattr(self, "first", 0):] or self.rows[-3:]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def global_triples(n_ranks, scaling):
    return BATCH if scaling == "strong" else BATCH * n_ranks


def build_problem(n_ranks, scaling="weak"):
    from neural_tensor_network_for_kb_completion_b200 import synth
    T = max(global_triples(n_ranks, scaling), BATCH)
    if SHAPE == "scaled":
        return synth.make_kb(SCALED["Ne"], SCALED["R"], T, SCALED["Nw"], SCALED["d"], seed=0)
    return synth.make_shaped(SHAPE, d=D, Nw=NW, T=T, seed=0)


def theta_seed(config):
    """c3: the seed recorded in the golden vector (the first one, from 11 up, whose float64 hinge margins stay clear of 0,
    tests/golden/make_c3_golden.py), so that the bench evaluates exactly the inputs the golden vector was made from."""
    if config == "c3" and os.path.exists(GOLDEN_C3):
        try:
            return int(np.load(GOLDEN_C3)["theta_seed"])
        except Exception:
            pass
    return THETA_SEED_DEFAULT


def theta_trained_like(theta0, seed=THETA_SEED_DEFAULT):
    """theta0 + N(0, 0.05^2): the hinge is inactive on part of the columns."""
    rng = np.random.default_rng(seed)
    return (theta0.astype(np.float64) + rng.standard_normal(theta0.size) * 0.05).astype(np.float32)


def initial_theta(kb):
    """NTN.java:67-74: W ~ U[0,1) * r with r = 1/sqrt(2d), V = 0.4, b = 0.01, U = 0.8, word block = the loaded vectors."""
    rng = np.random.default_rng(3)
    r0 = 1.0 / np.sqrt(2.0 * D)
    return np.concatenate([rng.random(kb.R * K_SLICES * D * D) * r0, np.full(kb.R * 2 * D * K_SLICES, 0.4),
                           np.full(kb.R * K_SLICES, 0.01), np.full(kb.R * K_SLICES, 0.8),
                           kb.wv.astype(np.float64).ravel()]).astype(np.float32)


def draw_sides(kb, rel_present, j):
    """corrupt_side[r] = JavaRandom(7 + j).nextDouble() > 0.5 for the relations present in the job (NTN.java:146)."""
    from neural_tensor_network_for_kb_completion_b200.data_factory import JavaRandom
    sr = JavaRandom(7 + j)
    present = np.zeros(kb.R, bool); present[np.unique(rel_present)] = True
    return np.array([1 if (present[r] and sr.nextDouble() > 0.5) else 0 for r in range(kb.R)], np.uint8)


def headline_problem():
    """The inputs of batch job 0 of the headline config at N=1 as oracle objects: (kb, NTNProblem, maps, batch, side,
    theta0).  Used by tests/golden/make_c3_golden.py and the golden GPU test: one definition of the workload."""
    from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory
    from oracle import ntn_ref
    global SHAPE, BATCH, CONFIG_NAME
    SHAPE, BATCH, CONFIG_NAME = CONFIGS["c3"]
    kb = build_problem(1)
    tbj = DataFactory.from_kb(kb, BATCH, CORRUPT, seed=42)
    maps = tbj.entityWordMaps()
    batch = tbj.generateNewTrainingBatchJob()
    side = draw_sides(kb, batch[1], 0)
    p = ntn_ref.NTNProblem(d=D, k=K_SLICES, R=kb.R, Ne=kb.Ne, Nw=kb.Nw, lam=LAMBDA, batch_size=BATCH,
                           fwd_off=maps[0], fwd_idx=maps[1], bwd_off=maps[2], bwd_idx=maps[3], len_bwd=maps[4])
    return kb, p, maps, batch, side, initial_theta(kb)


def check_against_golden(cost, grad, nact, theta_host):
    """This run's host-buffer result against the float64-oracle golden vector: every stored gradient entry and the cost
    within 1e-6 + 1e-4|ref| (BASELINE.json north_star), n_active exact."""
    try:
        g = np.load(GOLDEN_C3)
    except Exception as e:
        return {"status": "no golden vector", "detail": str(e)}
    chk = float(np.dot(theta_host[::37], np.arange(len(theta_host[::37])) % 7 + 1.0))
    if abs(chk - float(g["theta_checksum"])) > 1e-9 * max(1.0, abs(chk)):
        return {"status": "inputs differ from the golden vector's", "theta_checksum": chk}
    idx = g["idx"].astype(np.int64)
    ref = g["grad_at_idx"]
    err = np.abs(grad[idx] - ref)
    tol = 1e-6 + 1e-4 * np.abs(ref)
    offs = g["offsets"]
    blocks = [float(np.linalg.norm(grad[a:b])) for a, b in zip(offs[:-1], offs[1:])]
    blk_rel = [abs(x - y) / max(y, 1e-30) for x, y in zip(blocks, g["block_l2"])]
    ok = bool(np.all(err <= tol) and abs(cost - float(g["cost"])) <= 1e-6 + 1e-4 * abs(float(g["cost"])) and
              int(nact) == int(g["n_active"]))
    return {"status": "pass" if ok else "FAIL", "reference": "float64 oracle (oracle/ntn_ref.py), tests/golden/c3_fullsize.npz",
            "tolerance": "1e-6 + 1e-4*|ref|", "entries_checked": int(len(idx)), "entries_out_of_tolerance": int(np.sum(err > tol)),
            "grad_max_abs_diff": float(err.max()), "grad_max_tol_ratio": float(np.max(err / tol)),
            "cost": cost, "cost_ref": float(g["cost"]), "n_active": int(nact), "n_active_ref": int(g["n_active"]),
            "oracle_min_abs_margin": float(g["min_abs_margin"]), "block_l2_rel_diff_max": float(max(blk_rel))}


def work_model(dm, n_u, n_cols, nnz_f, nnz_b):
    """Algorithmic (compulsory) HBM bytes and FLOPs per kernel group for one evaluation: every
    buffer a group must read or write is counted ONCE -- row gathers that re-read a table (entity rows, T' rows)
    are served by the 126 MB L2 and are not counted, exactly like SURVEY.md §8d's B_alg."""
    d, k, R, Ne, Nw = dm["d"], dm["k"], dm["R"], dm["Ne"], dm["Nw"]
    P = R * k * d * d + R * 2 * d * k + 2 * R * k + d * Nw
    row, aug = d * 4, (d + 1) * (k * d + k)
    trow = 4 * (k * d + k)                                   # one T' / M row
    ents = min(Ne, 2 * n_u + n_cols)                         # distinct entity rows a batch can touch
    g = {
        "entity_build": dict(bytes=min(nnz_f, Nw) * row + Ne * row + 4 * (Ne + 1 + nnz_f), flops=nnz_f * d),
        "w_prep": dict(bytes=2 * 4 * R * aug, flops=0),
        "gemm_fwd": dict(bytes=min(n_u, Ne) * row + n_u * trow + 4 * R * aug, flops=2 * n_u * aug),
        "score": dict(bytes=2 * n_u * trow + ents * row + 4 * n_cols * (k + 1) + 12 * n_u, flops=4 * k * d * (n_u + n_cols)),
        "gemm_dw": dict(bytes=min(n_u, Ne) * row + n_u * trow + 4 * R * aug, flops=2 * n_u * aug),
        "gemm_ge": dict(bytes=n_u * trow + n_u * row + 4 * R * aug, flops=2 * n_u * aug),
        "entity_pull": dict(bytes=n_u * trow + (n_u + n_cols) * (4 * k + 16) + n_u * row + Ne * row,
                            flops=2 * k * d * (n_u + n_cols)),
        "word_pull": dict(bytes=Ne * row + 2 * Nw * row + 4 * (Nw + 1 + nnz_b), flops=3 * Nw * d),
        "finalize": dict(bytes=3 * 4 * R * aug, flops=R * aug),
    }
    b_alg = 8 * P + 16 * Ne * d + 12 * n_u + 4 * n_cols + 4 * (Ne + 1 + nnz_f)
    f_alg = 6 * d * d * k * n_u + 8 * d * k * n_cols
    return g, b_alg, f_alg, P


def workload_config(kb, n_ranks, wv_source, scaling):
    """The workload, and nothing measured: identical for the repo arm and `--impl reference`."""
    gt = global_triples(n_ranks, scaling)
    per = f"{BATCH} triples x {CORRUPT} corrupt per GPU" if scaling == "weak" else \
        f"one job of {BATCH} triples x {CORRUPT} corrupt sharded over the GPUs"
    return {"workload": f"{SHAPE.capitalize()}-shaped synthetic: {kb.Ne} entities, {kb.R} relations, d={D}, k={K_SLICES}, "
                        f"{per} ({CONFIG_NAME})",
            "entities": kb.Ne, "relations": kb.R, "d": D, "k": K_SLICES, "words": kb.Nw,
            "batch_per_gpu": gt // n_ranks, "corrupt": CORRUPT, "columns_per_gpu": gt // n_ranks * CORRUPT,
            "global_batch": gt, "wv_source": wv_source, "theta": "theta0 + N(0,0.05^2)",
            "parallelism": f"dp{n_ranks}" if n_ranks > 1 else "single", "scaling": scaling,
            "l2_policy": "rotate 2 batch jobs and 2 theta/grad buffers between steps; per-evaluation working set "
                         "(~190 MB) exceeds the 126 MB L2"}


def run_reference(args):
    """--impl reference: the oracle's float32 restatement of the reference formulation on the host
    cores.  The Java/ND4j reference cannot run here (no JVM, jars not vendored)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import data_ref, ntn_ref
    from oracle.java_random import JavaRandom
    n_ranks = max(1, args.gpus)
    gt = global_triples(n_ranks, args.scaling)
    kb = build_problem(n_ranks, args.scaling)
    vocab = {w: i for i, w in enumerate(kb.vocab_words)}
    fo, fi, bo, bi, lb = data_ref.build_entity_word_maps(kb.entity_names, vocab)
    p = ntn_ref.NTNProblem(d=D, k=K_SLICES, R=kb.R, Ne=kb.Ne, Nw=kb.Nw, lam=LAMBDA, batch_size=gt,
                           fwd_off=fo, fwd_idx=fi, bwd_off=bo, bwd_idx=bi, len_bwd=lb)
    e1, rel, e2, e3 = data_ref.generate_batch_job(kb.train[:, 0], kb.train[:, 1], kb.train[:, 2], gt, CORRUPT,
                                                  kb.Ne, JavaRandom(42))
    side = data_ref.draw_corrupt_sides(rel, kb.R, JavaRandom(7))
    theta = theta_trained_like(ntn_ref.initial_theta(p, kb.wv), theta_seed(args.config))

    def evaluate(m):
        """One evaluation of the first m training triples of the batch job with all their corrupt copies (h-major job:
        column j belongs to triple j mod gt); m == gt is the full job."""
        sel = slice(None) if m >= gt else (np.arange(len(e1)) % gt) < m
        t0 = time.perf_counter()
        ntn_ref.compute_at(theta, p, e1[sel], rel[sel], e2[sel], e3[sel], side, wv_frozen=None, dtype=np.float32)
        return time.perf_counter() - t0

    # calibrate on a 2,000-triple prefix, then size a step so that the whole run stays near `budget` seconds
    total_steps = args.steps + args.warmup
    t_cal = evaluate(min(2000, gt))
    rate = min(2000, gt) / t_cal                                   # triples/s, pessimistic (fixed costs dominate a small prefix)
    budget = 200.0
    m = int(min(gt, max(1000, rate * budget / max(total_steps, 1))))
    if m >= 0.9 * gt:
        m = gt
    for _ in range(args.warmup):
        evaluate(m)
    tt = 0.0
    for _ in range(args.steps):
        tt += evaluate(m)
    cores = os.cpu_count()
    val = m * args.steps / tt
    sample = (f"per step: ONE FULL evaluation of the {gt:,} x {CORRUPT} batch job" if m == gt else
              f"per step: one evaluation of the first {m:,} of the {gt:,} training triples of the batch job with all their "
              f"{CORRUPT} corrupt copies (a full evaluation does not fit {total_steps} steps in ~{budget:.0f} s)") + \
        ", float32 NumPy/OpenBLAS/SciPy restatement of the reference formulation"
    emit({
        "impl": "reference", "metric": f"NTN cost+grad triples/sec ({SHAPE.capitalize()} shape)", "value": val, "unit": "triples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(kb, n_ranks, "theta", args.scaling),
        "cpu_baseline": {"value": val, "unit": "triples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port of the reference formulation (NTN.java:92-443); the ND4j/JVM reference cannot run in this image",
    })


_REAL_STDOUT = None


def emit(obj):
    """The contract is ONE JSON line on stdout; libraries (NCCL's version banner) write there too,
    so fd 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""
    line = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--wv-source", default="theta", choices=["theta", "frozen"])
    ap.add_argument("--gemm-path", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline", action="store_true",
                    help="time the CPU port also on c4/c5 (one c4 evaluation is ~1 min of CPU; c5 is tens of minutes)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = the config's batch per GPU; strong = ONE job of the config's batch sharded over the GPUs")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "library"],
                    help="N > 1: fused = the library's own peer-memory exchange kernels inside the evaluation; "
                         "library = a NCCL / symmetric-memory all-reduce after it")
    ap.add_argument("--no-lbfgs", action="store_true")
    ap.add_argument("--batch", type=int, default=None,
                    help="override the config's batch-job size (training triples; c5: 10,000,000 = BASELINE.json configs[4] in full, "
                         "sharded with --scaling strong over the GPUs)")
    args = ap.parse_args()
    global SHAPE, BATCH, CONFIG_NAME, D, K_SLICES, NW
    SHAPE, BATCH, CONFIG_NAME = CONFIGS[args.config]
    if args.batch:
        BATCH = int(args.batch)
        CONFIG_NAME = CONFIG_NAME.split(",")[0] + f", {BATCH:,}-triple batch job"
    if SHAPE == "scaled":
        D, K_SLICES, NW = SCALED["d"], SCALED["k"], SCALED["Nw"]
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = 4000 if args.steps is None else args.steps        # ~0.9 s timed region at the headline config: nvidia-smi answers every 0.1-0.2 s
    args.warmup = max(3, 10 if args.warmup is None else args.warmup)

    import torch
    from neural_tensor_network_for_kb_completion_b200 import _capi
    from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    prio = -1 if os.environ.get("NTN_STREAM_PRIORITY", "1") != "0" else 0
    stream = torch.cuda.Stream(device=dev, priority=prio)   # everything (kernels, NCCL, timing events) on this stream; the
    #                                                         library's forked branches run on lowest-priority streams
    torch.cuda.set_stream(stream)

    kb = build_problem(world, args.scaling)
    gt = global_triples(world, args.scaling)
    hbm_peak, bf16_peak, peak_kind = peaks()
    tf32_peak = measure_tf32_tflops(dev)

    # two batch jobs (rotated between steps).  N > 1: ONE global batch job of `gt` triples (same
    # e3 stream on every rank), sharded by unique triple round-robin within each relation
    from neural_tensor_network_for_kb_completion_b200.distributed import shard_columns
    handles, sides = [], []
    set_batch_ms, set_batch_steady_ms = [], []
    local_triples = 0
    for j in range(2):
        tbj = DataFactory.from_kb(kb, gt, CORRUPT, seed=42 + 1000 * j)
        h = _capi.Handle(D, K_SLICES, kb.R, kb.Ne, kb.Nw, gt, LAMBDA, wv_source=args.wv_source,
                         gemm_path=args.gemm_path, device=local_rank, reg_scale=1.0 / world)
        maps = tbj.entityWordMaps()
        h.set_entity_words(*maps)
        if args.wv_source == "frozen":
            h.set_frozen_wv(kb.wv)
        bj = shard_columns(*tbj.generateNewTrainingBatchJob(), rank, world)
        local_triples = len(bj[0]) // CORRUPT
        t0 = time.perf_counter()
        h.set_batch(*bj)
        set_batch_ms.append(1e3 * (time.perf_counter() - t0))
        t0 = time.perf_counter()
        h.set_batch(*bj)                                    # again: buffers and workspaces exist now (what a training loop pays per job)
        set_batch_steady_ms.append(1e3 * (time.perf_counter() - t0))
        h.set_stream(stream.cuda_stream)
        handles.append(h)
        sides.append(draw_sides(kb, tbj.batchjob[1], j))     # from the GLOBAL job: identical on every rank
    h0 = handles[0]
    P, Pint = h0.P, h0.Pint
    theta_host = theta_trained_like(initial_theta(kb), theta_seed(args.config)).astype(np.float64)
    thetas = [torch.empty(Pint, dtype=torch.float32, device=dev) for _ in range(2)]
    exchange, exchange_us, exchange_kind = None, {}, None
    if dist is not None:
        from neural_tensor_network_for_kb_completion_b200.distributed import GradientExchange
        exchange = GradientExchange(Pint + 4, dev, n_buffers=2)
        grads = exchange.buffers
        for h in handles:
            h.set_grad_tail(True)
        fused = False
        os.environ.setdefault("NTN_EXCHANGE_TIMEOUT_MS", "10000")   # bound of the exchange's flag waits (graph capture skews the first evaluation)
        if args.exchange in ("auto", "fused"):
            fused = exchange.bind_fused(handles, grads)     # the library's own peer-memory exchange, inside the evaluation
            if args.exchange == "fused" and not fused:
                raise RuntimeError("fused exchange requested but symmetric memory is not available")
        if fused:
            exchange_kind = "fused (peer-memory reduce + broadcast kernels of libntn_b200.so, overlapped with the word pull)"
        else:
            exchange_us = exchange.autotune()
            exchange_kind = exchange.variant
    else:
        fused = False
        grads = [torch.empty(Pint + 4, dtype=torch.float32, device=dev) for _ in range(2)]
    for t in thetas:
        h0.upload_theta(theta_host, t.data_ptr())
    torch.cuda.synchronize()

    def step(i, sync=False):
        j = i & 1
        h = handles[j]
        h.eval_dev(thetas[j].data_ptr(), sides[j], grads[j].data_ptr(), sync=False)
        if dist is not None and not fused:
            exchange.all_reduce(grads[j])                   # every rank ends with the full-batch gradient, cost, n_active
        if sync:
            torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i, sync=True)
    if fused:
        for h in handles:
            h.exchange_check()          # (clears the flag: a rank that was late to its FIRST evaluation does not void the run)
    launches_per_eval = h0.launches
    barrier()
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step(i)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    cost, nact = handles[(args.steps - 1) & 1].last_cost()
    if fused and any(h.exchange_check() for h in handles):
        raise RuntimeError("gradient exchange: a flag wait expired during the timed region (a rank fell > 2 s behind); results invalid")
    if dist is not None:
        cost, nact = _capi.Handle.unpack_tail(grads[(args.steps - 1) & 1][Pint:Pint + 4].cpu().numpy())
    # per-kernel-group times: the same steps again with CUDA events bracketing every group on the stream it
    # runs on (this mode launches the kernels directly instead of replaying the captured graph)
    for h in handles:
        h.set_profiling(True)
    for i in range(min(args.steps, 40)):
        step(i)
    barrier()
    phase = {}
    for h in handles:
        for kname, v in h.phase_ms().items():
            phase[kname] = phase.get(kname, 0.0) + v / len(handles)
        h.set_profiling(False)
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    triples_per_s = gt / (ms_step * 1e-3)

    # ---- end to end through the reference-facing call with host buffers ---------------------
    grad_host = np.empty(P, np.float64)
    e2e_steps = max(1, args.e2e_steps)

    def e2e_step(i):
        j = i & 1
        if dist is None:
            return handles[j].eval(theta_host, sides[j], grad_out=grad_host)
        # the caller lives on rank 0: its host theta goes up once, travels to the other GPUs over NVLink, every rank
        # evaluates its shard, the gradient is exchanged and rank 0 alone brings it back to the host
        hj = handles[j]
        if rank == 0:
            hj.upload_theta(theta_host, thetas[j].data_ptr())
        dist.broadcast(thetas[j], src=0)
        hj.eval_dev(thetas[j].data_ptr(), sides[j], grads[j].data_ptr(), sync=False)
        if not fused:
            exchange.all_reduce(grads[j])
        if rank == 0:
            hj.download_theta(grads[j].data_ptr(), out=grad_host)
        return None
    last = None
    for i in range(3):
        last = e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        last = e2e_step(i)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # ---- parity of this run against the golden vector (headline config, N=1) ---------------------------------
    parity = None
    if rank == 0 and world == 1 and args.config == "c3" and args.wv_source == "theta":
        c0, g0, n0 = handles[0].eval(theta_host, sides[0])
        parity = check_against_golden(c0, g0, n0, theta_host)

    # ---- device L-BFGS on the same workload (Run_NTN.java:119-124): one minimize(maxIters=5) call ------------
    lbfgs = None
    if rank == 0 and world == 1 and not args.no_lbfgs and args.config in ("c1", "c3"):
        d_t = thetas[0].clone()
        torch.cuda.synchronize()
        hl = handles[0]
        res = hl.lbfgs_minimize(d_t.data_ptr(), lambda: sides[0], max_iters=5)       # warm-up (graph capture, workspace)
        h0.upload_theta(theta_host, d_t.data_ptr())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = hl.lbfgs_minimize(d_t.data_ptr(), lambda: sides[0], max_iters=5)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        m_hist = 6
        lbfgs_bytes = (4 * m_hist + 6) * 4 * P                       # (4m + 6) passes over theta per iteration
        vec_ms = max(1e3 * dt - res.evals * ms_step, 1e-6)
        lbfgs = {"ms_per_minimize": 1e3 * dt, "iters": int(res.iters), "evals": int(res.evals),
                 "ms_per_eval_all_in": 1e3 * dt / max(res.evals, 1), "vector_ops_ms_per_iter": vec_ms / max(res.iters, 1),
                 "bytes_per_iter_model": lbfgs_bytes,
                 "vector_ops_GBps": lbfgs_bytes * max(res.iters, 1) / (vec_ms * 1e-3) / 1e9,
                 "frac_of_hbm_peak": lbfgs_bytes * max(res.iters, 1) / (vec_ms * 1e-3) / 1e9 / hbm_peak,
                 "first_cost": res.first_cost, "min_obj_val": res.min_obj_val,
                 "note": "one ntn_lbfgs_minimize(maxIters=5) on batch job 0, host wall clock; vector-op time = total - evals x ms_per_step"}

    # ---- roofline of the slowest kernel group -------------------------------------------------
    dm = dict(d=D, k=K_SLICES, R=kb.R, Ne=kb.Ne, Nw=kb.Nw)
    nnz_f = int(maps[0][-1]); nnz_b = int(maps[2][-1])
    groups, b_alg, f_alg, _ = work_model(dm, local_triples, local_triples * CORRUPT, nnz_f, nnz_b)
    # dominant kernel group = the longest bracket on the MAIN chain.  The forked branches (weight prep, dW contraction,
    # finalize) run beside it and their event brackets include the time they spend sharing the SMs with the chain, so
    # they are reported under "phase" but never chosen here.
    side_branch = ("w_prep", "gemm_dw", "finalize")
    chain = {k_: v for k_, v in phase.items() if k_ not in side_branch}
    top = max(chain, key=chain.get) if chain else None
    tensor_groups = ("gemm_fwd", "gemm_dw", "gemm_ge")
    tensor_peak = tf32_peak / 3.0                                    # 3xTF32: three kind::tf32 MMAs per FP32-equivalent MAC
    roofline = None
    if top and phase[top] > 0:
        tsec = phase[top] * 1e-3
        if top in tensor_groups:
            ach = groups[top]["flops"] / tsec / 1e12
            roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s",
                        "frac": ach / tensor_peak, "traffic": None,
                        "peak_note": "measured in this run: cuBLAS kind::tf32 GEMM 8192^3 / 3 (3xTF32 split)"}
        else:
            ach = groups[top]["bytes"] / tsec / 1e9
            roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": None, "peak_note": f"{peak_kind} copy bandwidth"}
        for fn in ("r2_traffic.json", "r1_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", fn)) as f:
                    tr = json.load(f)
                if args.config == "c3" and world == 1 and top in tr:
                    roofline["traffic"] = tr[top]["dram_bytes_per_eval"]
                    roofline["traffic_note"] = (f"dram__bytes_read+write of this group's kernels, one evaluation, ncu --set full "
                                                f"(profiles/{fn}; N=1, same workload)")
                    break
            except Exception:
                pass
        roofline["algorithmic_bytes"] = groups[top]["bytes"]
        roofline["algorithmic_flops"] = groups[top]["flops"]
        roofline["ms"] = phase[top]
        roofline["share_of_step"] = phase[top] / max(sum(phase.values()), 1e-9)
    per_group = {kname: {"ms": v, "GB/s": groups[kname]["bytes"] / (v * 1e-3) / 1e9 if v > 0 else None,
                         "TFLOP/s": groups[kname]["flops"] / (v * 1e-3) / 1e12 if v > 0 else None}
                 for kname, v in phase.items()}
    tstep = ms_step * 1e-3
    eval_roofline = {"B_alg_bytes": b_alg, "F_alg_flops": f_alg,
                     "hbm": {"achieved": b_alg / tstep / 1e9, "peak": hbm_peak, "frac": b_alg / tstep / 1e9 / hbm_peak, "unit": "GB/s"},
                     "tensor": {"achieved": f_alg / tstep / 1e12, "peak": tensor_peak, "frac": f_alg / tstep / 1e12 / tensor_peak,
                                "unit": "TFLOP/s"},
                     "tf32_dense_tflops_measured": tf32_peak, "bf16_dense_tflops_measured": bf16_peak,
                     "note": "per rank; B_alg, F_alg per SURVEY.md §8d; HBM peak " + peak_kind + ", tensor peak = measured kind::tf32 / 3"}

    # ---- CPU baseline (rank 0, N=1 only): one full evaluation by the oracle port ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and (args.config in ("c1", "c3") or args.cpu_baseline):
        from oracle import ntn_ref                                  # checker / baseline only
        p = ntn_ref.NTNProblem(d=D, k=K_SLICES, R=kb.R, Ne=kb.Ne, Nw=kb.Nw, lam=LAMBDA, batch_size=BATCH,
                               fwd_off=maps[0], fwd_idx=maps[1], bwd_off=maps[2], bwd_idx=maps[3], len_bwd=maps[4])
        tbj = DataFactory.from_kb(kb, BATCH, CORRUPT, seed=42)
        bj = tbj.generateNewTrainingBatchJob()
        t0 = time.perf_counter()
        c_cpu, g_cpu = ntn_ref.compute_at(theta_host, p, *bj, sides[0], dtype=np.float32)
        dt = time.perf_counter() - t0
        cpu = {"value": BATCH / dt, "unit": "triples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 full evaluation of the {BATCH:,} x {CORRUPT} batch, float32 NumPy/OpenBLAS/SciPy restatement of "
                         "the reference formulation (oracle/ntn_ref.py); the ND4j/JVM reference cannot run here",
               "seconds_per_eval": dt}

    if rank == 0:
        out = {
            "metric": f"NTN cost+grad triples/sec ({SHAPE.capitalize()} shape)", "value": triples_per_s, "unit": "triples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(kb, world, args.wv_source, args.scaling),
            "runtime": {"gemm_path": h0.gemm_path, "set_batch_ms_first": float(np.mean(set_batch_ms)),
                        "set_batch_ms_steady": float(np.mean(set_batch_steady_ms)),
                        "exchange": exchange_kind, "exchange_us": exchange_us},
            "amortised_ms_per_step": ms_step + float(np.mean(set_batch_steady_ms)) / EVALS_PER_JOB,
            "amortised_note": f"ms_per_step + steady ntn_set_batch / {EVALS_PER_JOB} evaluations per batch job (one L-BFGS(5) call)",
            "evals_per_s": 1e3 / ms_step, "columns_per_s": gt * CORRUPT / (ms_step * 1e-3),
            "cost": cost, "n_active": nact,
            "clocks": clocks,
            "e2e": {"value": gt / e2e_s, "unit": "triples/s", "ms_per_step": 1e3 * e2e_s,
                    "h2d_bytes_per_step": 4 * P + kb.R, "d2h_bytes_per_step": 4 * P + 32,
                    "api": "ntn_eval(double* theta, uint8* corrupt_side, double* cost, double* grad) with host buffers"
                           if world == 1 else "rank 0: ntn_upload_theta(host double[]) -> NVLink broadcast -> ntn_eval_dev on every rank -> "
                                              "gradient exchange -> rank 0: ntn_download_theta(host double[])"},
            "gpu_launches": launches_per_eval * args.steps,
            "launches_per_eval": launches_per_eval,
            "roofline": roofline, "phase": per_group, "eval_roofline": eval_roofline,
            "parity": parity, "lbfgs": lbfgs,
            "cpu_baseline": cpu,
        }
        emit(out)
    for h in handles:
        h.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
