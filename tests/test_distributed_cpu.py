"""World-size-2 gloo test (CPU) of the data-parallel host logic: sharding by unique triple within
relation, reg_scale = 1/G, one all-reduce -> identical to the single-process evaluation.
The local evaluator is the oracle here (no GPU on this box); on GPUs it is the CUDA handle
(tests/test_gpu_parity.py::test_reg_scale_makes_rank_results_additive, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_tensor_network_for_kb_completion_b200.distributed import DataParallelNTN, shard_columns
from oracle import ntn_ref
from tests.helpers import perturbed, small_problem, spread


def test_shards_partition_the_batch_and_balance_relations():
    _, p, batch, _, _ = small_problem(Ne=200, R=5, T=900, Nw=100, d=8, batch=600, corrupt=4, seed=3)
    n = len(batch[0])
    seen = np.zeros(n, int)
    world = 4
    cols = np.stack(batch, 1)
    per_rank_rel = []
    for r in range(world):
        sh = np.stack(shard_columns(*batch, r, world), 1)
        # every triple of a shard carries all its corrupt copies
        key = {tuple(x[:3]) for x in sh}
        full = np.array([tuple(x[:3]) in key for x in cols])
        assert full.sum() == len(sh)
        seen += full
        per_rank_rel.append(np.bincount(sh[:, 1], minlength=p.R))
    assert np.all(seen == 1)                                   # disjoint cover
    spread_ = np.ptp(np.stack(per_rank_rel), axis=0)
    assert spread_.max() <= 4 * 1 + 4                          # <= one triple (x C copies, + duplicates) per relation


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, p, batch, side, theta0 = small_problem(Ne=200, R=5, T=900, Nw=100, d=8, batch=600, corrupt=4, seed=3)
    theta = spread(perturbed(theta0, 0.05), p)
    shard = shard_columns(*batch, rank, world)

    def local_eval(th, sd):
        c, g, aux = ntn_ref.compute_at(th, p, *shard, sd, return_aux=True)
        # emulate the handle's reg_scale = 1/G: the oracle applies the full regulariser
        reg_c, reg_g = 0.5 * p.lam * np.sum(th * th), p.lam * th
        c = c - reg_c + reg_c / world
        g = g - reg_g + reg_g / world
        return c, torch.from_numpy(g), aux["n_active"]

    dp = DataParallelNTN(local_eval, p.R, rank, world, draw_sides=lambda: side)
    cost, grad, nact = dp.computeAt(theta)
    if rank == 0:
        c_ref, g_ref, aux = ntn_ref.compute_at(theta, p, *batch, side, return_aux=True)
        out.put((abs(cost - c_ref) / abs(c_ref), float(np.max(np.abs(grad.numpy() - g_ref))), nact, aux["n_active"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    mp.spawn(_worker, args=(2, port, q), nprocs=2, join=True)
    rel_c, max_g, nact, nact_ref = q.get()
    assert rel_c < 1e-12 and max_g < 1e-12 and nact == nact_ref
