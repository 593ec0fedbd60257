"""Data-parallel training check (run under torchrun on >= 2 GPUs): the Run_NTN loop (new batch job -> L-BFGS with
maxIters 5, Run_NTN.java:113-137) with the batch job sharded over the ranks and one all-reduce per evaluation must
follow the single-GPU run on the same global batch, and all ranks must hold bit-identical theta.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/dp_train_check.py [wordnet]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from neural_tensor_network_for_kb_completion_b200 import synth
from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory
from neural_tensor_network_for_kb_completion_b200.distributed import GradientExchange, make_lbfgs_reduce, shard_columns
from neural_tensor_network_for_kb_completion_b200.lbfgs import LBFGSMinimizer, Opts
from neural_tensor_network_for_kb_completion_b200.ntn import NTN

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
big = len(sys.argv) > 1 and sys.argv[1] == "wordnet"
if big:
    kb = synth.make_shaped("wordnet", d=100, Nw=67447, T=20000, seed=0); B, C, k, outer = 20000, 10, 3, 6
else:
    kb = synth.make_kb(Ne=3000, R=6, T=4000, Nw=2500, d=36, seed=5); B, C, k, outer = 4000, 4, 3, 3


def make(reg_scale, dev):
    tbj = DataFactory.from_kb(kb, B, C, seed=42)                       # same seed -> same global batch jobs on every rank
    t = NTN(kb.d, kb.Ne, kb.R, kb.Nw, B, k, "tanh", tbj, 1e-4, wv_source="frozen", device=dev, reg_scale=reg_scale)
    return tbj, t


tbj, t = make(1.0 / world, local)
stream = torch.cuda.Stream()
t.handle.set_stream(stream.cuda_stream)
ex = GradientExchange(t.handle.Pint + 4, torch.device(f"cuda:{local}"))
tune = ex.autotune()
reduce = make_lbfgs_reduce(ex, torch.device(f"cuda:{local}"), stream)
theta = np.asarray(t.getTheta_inital(), np.float64)
if rank == 0:
    tbj1, t1 = make(1.0, local)
    theta1 = theta.copy()
    print(f"world {world}, all-reduce variants (us): {tune} -> {ex.variant}", flush=True)
ok = True
for it in range(outer):
    job = tbj.generateNewTrainingBatchJob()
    t.global_relations = job[1]
    tbj.batchjob = shard_columns(*job, rank, world)
    t._sync_batch(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    res = LBFGSMinimizer().minimize(t, theta, Opts(maxIters=5), reduce=reduce)
    torch.cuda.synchronize(); ms = 1e3 * (time.perf_counter() - t0)
    theta = res.minArg
    # lock-step: every rank's theta bit-identical
    h = torch.tensor([float(np.sum(theta)), float(np.abs(theta).sum())], dtype=torch.float64, device="cuda")
    lo, hi = h.clone(), h.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((lo == hi).all())
    if rank == 0:
        tbj1.generateNewTrainingBatchJob()
        t0 = time.perf_counter()
        r1 = LBFGSMinimizer().minimize(t1, theta1, Opts(maxIters=5)); theta1 = r1.minArg
        ms1 = 1e3 * (time.perf_counter() - t0)
        dth = float(np.max(np.abs(theta - theta1)) / (np.max(np.abs(theta1)) + 1e-30))
        dc = abs(res.minObjVal - r1.minObjVal) / abs(r1.minObjVal)
        good = same and res.evals == r1.evals and dc < 1e-4 and dth < 1e-3
        ok &= good
        print(f"outer {it}: DP cost {res.firstObjVal:.6f} -> {res.minObjVal:.6f} ({res.evals} evals, {ms:.1f} ms) | 1 GPU "
              f"{r1.firstObjVal:.6f} -> {r1.minObjVal:.6f} ({r1.evals} evals, {ms1:.1f} ms) | rel cost diff {dc:.2e}, "
              f"theta diff {dth:.2e}, ranks identical {same} {'OK' if good else 'MISMATCH'}", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.broadcast(flag, src=0)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DP_TRAIN_CHECK", "PASS" if ok else "FAIL", flush=True)
sys.exit(0 if flag.item() == 1.0 else 1)
