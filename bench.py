#!/usr/bin/env python
"""bench.py -- NTN cost+gradient evaluations on Freebase-shaped synthetic data (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one cost+gradient evaluation (NTN.computeAt, /root/reference/src/NTN.java:92-443)
of one batch job: 20,000 training triples x 10 corrupt copies per GPU (Run_NTN.java:72-77) on
the Freebase shape (75,043 entities, 13 relations, d=100, k=3; BASELINE.json configs[2]).

  value        triples/s, whole job, theta and batch already resident in HBM (ntn_eval_dev)
  e2e          the same through the reference-facing call ntn_eval(double* theta, ..., double* grad)
               with HOST buffers: staging, H2D, layout conversion, evaluation, D2H inside the timer
  roofline     the slowest kernel group of the evaluation, timed with CUDA events on the stream
  cpu_baseline the oracle's float32 restatement of the reference formulation on the host cores
               (N=1, rank 0): one full evaluation

N > 1 (torchrun): weak scaling -- every rank evaluates its own 20,000 x 10 shard of a G*20,000
triple batch and the gradient is all-reduced (NCCL) so every rank holds the full-batch gradient.
`--impl reference` times the oracle port alone (rank 0 only).
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
           "c4": ("freebase", 316232, "BASELINE.json configs[3], per-GPU share of the full batch"),
           # configs[4]'s model shape (1M entities, 64 relations, d=256, k=8, 500k words) with a 200,000 x 10 batch job
           "c5": ("scaled", 200000, "BASELINE.json configs[4] shape, 200,000-triple batch job")}
SCALED = dict(Ne=1000000, R=64, d=256, k=8, Nw=500000)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        return float(m["hbm_gbs"]), float(m["bf16_tflops_sustained"]), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


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


def build_problem(n_ranks):
    from neural_tensor_network_for_kb_completion_b200 import synth
    if SHAPE == "scaled":
        return synth.make_kb(SCALED["Ne"], SCALED["R"], max(BATCH * n_ranks, BATCH), SCALED["Nw"], SCALED["d"], seed=0)
    kb = synth.make_shaped(SHAPE, d=D, Nw=NW, T=max(BATCH * n_ranks, BATCH), seed=0)
    return kb


def theta_trained_like(theta0, seed=11):
    """theta0 + N(0, 0.05^2): the hinge is inactive on part of the columns."""
    rng = np.random.default_rng(seed)
    return (theta0.astype(np.float64) + rng.standard_normal(theta0.size) * 0.05).astype(np.float32)


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


def run_reference(args):
    """--impl reference: the oracle's float32 restatement of the reference formulation on the host
    cores.  The Java/ND4j reference cannot run here (no JVM, jars not vendored)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import data_ref, ntn_ref
    from oracle.java_random import JavaRandom
    kb = build_problem(1)
    vocab = {w: i for i, w in enumerate(kb.vocab_words)}
    fo, fi, bo, bi, lb = data_ref.build_entity_word_maps(kb.entity_names, vocab)
    p = ntn_ref.NTNProblem(d=D, k=K_SLICES, R=kb.R, Ne=kb.Ne, Nw=kb.Nw, lam=LAMBDA, batch_size=BATCH,
                           fwd_off=fo, fwd_idx=fi, bwd_off=bo, bwd_idx=bi, len_bwd=lb)
    e1, rel, e2, e3 = data_ref.generate_batch_job(kb.train[:, 0], kb.train[:, 1], kb.train[:, 2], BATCH, CORRUPT,
                                                  kb.Ne, JavaRandom(42))
    side = data_ref.draw_corrupt_sides(rel, kb.R, JavaRandom(7))
    theta = theta_trained_like(ntn_ref.initial_theta(p, kb.wv))
    # bounded sample: the columns of m relations per step, rotating, m sized for ~200 s overall
    total_steps = args.steps + args.warmup
    m = int(max(1, min(kb.R, (200.0 / total_steps - 0.5) / 1.1)))
    rel_sets = [[(i * m + j) % kb.R for j in range(m)] for i in range(total_steps)]

    def step(i):
        sel = np.isin(rel, rel_sets[i])
        t0 = time.perf_counter()
        ntn_ref.compute_at(theta, p, e1[sel], rel[sel], e2[sel], e3[sel], side, wv_frozen=None, dtype=np.float32)
        return time.perf_counter() - t0, int(sel.sum()) // CORRUPT

    for i in range(args.warmup):
        step(i)
    tt, trip = 0.0, 0
    for i in range(args.warmup, total_steps):
        dt, n = step(i)
        tt += dt; trip += n
    cores = os.cpu_count()
    val = trip / tt
    sample = f"per step: all columns of {m} of {kb.R} relations (rotating) of the 20,000x10 batch, float32 NumPy/OpenBLAS/SciPy"
    emit({
        "impl": "reference", "metric": "NTN cost+grad triples/sec (Freebase shape)", "value": val, "unit": "triples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(kb, 1, "theta"),
        "cpu_baseline": {"value": val, "unit": "triples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port of the reference formulation (NTN.java:92-443); the ND4j/JVM reference cannot run in this image",
    })


def workload_config(kb, n_ranks, wv_source):
    return {"workload": f"{SHAPE.capitalize()}-shaped synthetic: {kb.Ne} entities, {kb.R} relations, d={D}, k={K_SLICES}, "
                        f"{BATCH} triples x {CORRUPT} corrupt per GPU ({CONFIG_NAME})",
            "entities": kb.Ne, "relations": kb.R, "d": D, "k": K_SLICES, "words": kb.Nw,
            "batch_per_gpu": BATCH, "corrupt": CORRUPT, "columns_per_gpu": BATCH * CORRUPT,
            "global_batch": BATCH * n_ranks, "wv_source": wv_source, "theta": "theta0 + N(0,0.05^2)",
            "parallelism": f"dp{n_ranks}" if n_ranks > 1 else "single",
            "l2_policy": "rotate 2 batch jobs and 2 theta/grad buffers between steps; per-evaluation working set "
                         "(~190 MB) exceeds the 126 MB L2"}


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
    args = ap.parse_args()
    global SHAPE, BATCH, CONFIG_NAME, D, K_SLICES, NW
    SHAPE, BATCH, CONFIG_NAME = CONFIGS[args.config]
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
    from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory, JavaRandom

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

    kb = build_problem(world)
    hbm_peak, tf_peak, peak_kind = peaks()

    # two batch jobs (rotated between steps).  N > 1: ONE global batch job of world*BATCH triples (same
    # e3 stream on every rank), sharded by unique triple round-robin within each relation
    from neural_tensor_network_for_kb_completion_b200.distributed import shard_columns
    handles, sides, results = [], [], []
    set_batch_ms = []
    for j in range(2):
        tbj = DataFactory.from_kb(kb, BATCH * world, CORRUPT, seed=42 + 1000 * j)
        h = _capi.Handle(D, K_SLICES, kb.R, kb.Ne, kb.Nw, BATCH * world, LAMBDA, wv_source=args.wv_source,
                         gemm_path=args.gemm_path, device=local_rank, reg_scale=1.0 / world)
        maps = tbj.entityWordMaps()
        h.set_entity_words(*maps)
        if args.wv_source == "frozen":
            h.set_frozen_wv(kb.wv)
        bj = shard_columns(*tbj.generateNewTrainingBatchJob(), rank, world)
        t0 = time.perf_counter()
        h.set_batch(*bj)
        set_batch_ms.append(1e3 * (time.perf_counter() - t0))
        h.set_stream(stream.cuda_stream)
        handles.append(h)
        results.append(h.result_tensor())
        sr = JavaRandom(7 + j)
        present = np.zeros(kb.R, bool); present[np.unique(bj[1])] = True
        sides.append(np.array([1 if (present[r] and sr.nextDouble() > 0.5) else 0 for r in range(kb.R)], np.uint8))
    h0 = handles[0]
    P, Pint = h0.P, h0.Pint
    rng = np.random.default_rng(3)
    r0 = 1.0 / np.sqrt(2.0 * D)
    theta0 = np.concatenate([rng.random(kb.R * K_SLICES * D * D) * r0, np.full(kb.R * 2 * D * K_SLICES, 0.4),
                             np.full(kb.R * K_SLICES, 0.01), np.full(kb.R * K_SLICES, 0.8),
                             kb.wv.astype(np.float64).ravel()]).astype(np.float32)
    theta_host = theta_trained_like(theta0).astype(np.float64)
    thetas = [torch.empty(Pint, dtype=torch.float32, device=dev) for _ in range(2)]
    exchange, exchange_us = None, {}
    if dist is not None:
        # one collective per evaluation: [gradient | cost, n_active] in one buffer, fastest available all-reduce
        from neural_tensor_network_for_kb_completion_b200.distributed import GradientExchange
        exchange = GradientExchange(Pint + 4, dev, n_buffers=2)
        exchange_us = exchange.autotune()
        grads = exchange.buffers
        for h in handles:
            h.set_grad_tail(True)
    else:
        grads = [torch.empty(Pint + 4, dtype=torch.float32, device=dev) for _ in range(2)]
    for t in thetas:
        h0.upload_theta(theta_host, t.data_ptr())
    torch.cuda.synchronize()

    def step(i, sync=False):
        j = i & 1
        h = handles[j]
        h.eval_dev(thetas[j].data_ptr(), sides[j], grads[j].data_ptr(), sync=False)
        if dist is not None:
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
    cost, nact = handles[(args.steps - 1) & 1].last_cost()
    if dist is not None:
        cost, nact = _capi.Handle.unpack_tail(grads[(args.steps - 1) & 1][Pint:Pint + 4].cpu().numpy())
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    triples_per_s = BATCH * world / (ms_step * 1e-3)

    # ---- end to end through the reference-facing call with host buffers ---------------------
    grad_host = np.empty(P, np.float64)
    e2e_steps = max(1, args.e2e_steps)

    def e2e_step(i):
        j = i & 1
        if dist is None:
            handles[j].eval(theta_host, sides[j], grad_out=grad_host)
        else:
            hj = handles[j]
            hj.upload_theta(theta_host, thetas[j].data_ptr())
            hj.eval_dev(thetas[j].data_ptr(), sides[j], grads[j].data_ptr(), sync=False)
            exchange.all_reduce(grads[j])
            hj.download_theta(grads[j].data_ptr(), out=grad_host)
    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # ---- roofline of the slowest kernel group -------------------------------------------------
    dm = dict(d=D, k=K_SLICES, R=kb.R, Ne=kb.Ne, Nw=kb.Nw)
    nnz_f = int(maps[0][-1]); nnz_b = int(maps[2][-1])
    groups, b_alg, f_alg, _ = work_model(dm, BATCH, BATCH * CORRUPT, nnz_f, nnz_b)
    # dominant kernel group = the longest bracket on the MAIN chain.  The forked branches (weight prep, dW contraction,
    # finalize) run beside it and their event brackets include the time they spend sharing the SMs with the chain, so
    # they are reported under "phase" but never chosen here.
    side_branch = ("w_prep", "gemm_dw", "finalize")
    chain = {k_: v for k_, v in phase.items() if k_ not in side_branch}
    top = max(chain, key=chain.get) if chain else None
    tensor_groups = ("gemm_fwd", "gemm_dw", "gemm_ge")
    roofline = None
    if top and phase[top] > 0:
        tsec = phase[top] * 1e-3
        if top in tensor_groups:
            # 3xTF32 on tcgen05: FP32-equivalent peak = (bf16/2)/3 (no TF32 figure is measured)
            pk = tf_peak / 6.0
            ach = groups[top]["flops"] / tsec / 1e12
            roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s",
                        "frac": ach / pk, "traffic": None,
                        "peak_note": f"{peak_kind} bf16 sustained / 2 (tf32) / 3 (3xTF32 split)"}
        else:
            ach = groups[top]["bytes"] / tsec / 1e9
            roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": None, "peak_note": f"{peak_kind} copy bandwidth"}
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tr = json.load(f)
            if args.config == "c3" and top in tr:
                roofline["traffic"] = tr[top]["dram_bytes_per_eval"]
                roofline["traffic_note"] = ("dram__bytes_read+write of this group's kernels, one evaluation, from the ncu --set full "
                                            "capture in profiles/r1_ncu_full_summary.txt (N=1, same workload)")
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
                     "tensor": {"achieved": f_alg / tstep / 1e12, "peak": tf_peak / 6.0, "frac": f_alg / tstep / 1e12 / (tf_peak / 6.0), "unit": "TFLOP/s"},
                     "note": "B_alg, F_alg per SURVEY.md §8d; peaks " + peak_kind}

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
        c_gpu, g_gpu, _ = handles[0].eval(theta_host, sides[0])
        cpu = {"value": BATCH / dt, "unit": "triples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 full evaluation of the {BATCH:,} x {CORRUPT} batch, float32 NumPy/OpenBLAS/SciPy restatement of "
                         "the reference formulation (oracle/ntn_ref.py); the ND4j/JVM reference cannot run here",
               "seconds_per_eval": dt,
               "gpu_vs_cpu_f32_cost_rel_diff": abs(c_gpu - c_cpu) / max(abs(c_cpu), 1e-30),
               "gpu_vs_cpu_f32_grad_max_abs_diff": float(np.max(np.abs(g_gpu - g_cpu)))}

    if rank == 0:
        out = {
            "metric": f"NTN cost+grad triples/sec ({SHAPE.capitalize()} shape)", "value": triples_per_s, "unit": "triples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(kb, world, args.wv_source), gemm_path=h0.gemm_path,
                           set_batch_ms=float(np.mean(set_batch_ms)),
                           exchange=(exchange.variant if exchange else None), exchange_us=exchange_us),
            "evals_per_s": 1e3 / ms_step, "columns_per_s": BATCH * CORRUPT * world / (ms_step * 1e-3),
            "cost": cost, "n_active": nact,
            "clocks": clocks,
            "e2e": {"value": BATCH * world / e2e_s, "unit": "triples/s", "ms_per_step": 1e3 * e2e_s,
                    "h2d_bytes_per_step": 4 * P + kb.R, "d2h_bytes_per_step": 4 * P + 32,
                    "api": "ntn_eval(double* theta, uint8* corrupt_side, double* cost, double* grad) with host buffers"
                           if world == 1 else "ntn_upload_theta + ntn_eval_dev + gradient all-reduce + ntn_download_theta"},
            "gpu_launches": launches_per_eval * args.steps,
            "launches_per_eval": launches_per_eval,
            "roofline": roofline, "phase": per_group, "eval_roofline": eval_roofline,
            "cpu_baseline": cpu,
        }
        emit(out)
    for h in handles:
        h.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
