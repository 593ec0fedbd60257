"""Data-parallel evaluation of NTN.computeAt over the GPUs of one box.

Every batch column contributes independently to cost and gradient given theta
(/root/reference/src/NTN.java:114-409 is a sum over columns, :411-430 is linear in the entity
gradient), so the batch job is sharded and the per-rank results are summed:

  * unique training triples are dealt round-robin *within each relation* (every rank gets
    n_r/G triples of every relation -> equal grouped-contraction shapes), each triple keeps all of
    its corrupt copies (the de-duplication of DESIGN.md §3 survives);
  * every rank evaluates its shard with reg_scale = 1/G and batch_divisor = the GLOBAL batch size,
    so that ONE all-reduce(SUM) of [gradient | cost | n_active] is the full-batch result on every
    rank (the regulariser is counted G times at weight 1/G);
  * corrupt_side[r] is drawn once (rank 0's stream) and broadcast, like the reference's single
    Math.random() per relation per evaluation (NTN.java:146).

One process per GPU; torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU
tests).  The local evaluator is injected, so the same class drives the CUDA handle and the test's
CPU evaluator.
"""
from __future__ import annotations

import numpy as np


def shard_columns(e1, rel, e2, e3, rank: int, world: int):
    """Columns of this rank: unique (rel, e1, e2) triples round-robin within their relation, in
    order of first appearance, all corrupt copies of a triple on the same rank."""
    e1, rel, e2, e3 = (np.asarray(x) for x in (e1, rel, e2, e3))
    if world == 1:
        return e1, rel, e2, e3
    key = (rel.astype(np.int64) << 42) | (e1.astype(np.int64) << 21) | e2.astype(np.int64)
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    # order unique triples by first appearance inside their relation
    u_rel = rel[first]
    order = np.lexsort((first, u_rel))
    pos_in_rel = np.empty(len(uniq), np.int64)
    sorted_rel = u_rel[order]
    starts = np.r_[0, np.nonzero(np.diff(sorted_rel))[0] + 1]
    start_of = np.repeat(starts, np.diff(np.r_[starts, len(order)]))
    pos_in_rel[order] = np.arange(len(order)) - start_of
    sel = (pos_in_rel[inv] % world) == rank
    return e1[sel], rel[sel], e2[sel], e3[sel]


class DataParallelNTN:
    """computeAt over `world` ranks.  ``local_eval(theta, corrupt_side) -> (cost, grad, n_active)``
    evaluates this rank's shard (grad: torch tensor on the rank's device, already including the
    1/G-scaled regulariser)."""

    def __init__(self, local_eval, num_relations: int, rank: int, world: int, draw_sides=None):
        import torch.distributed as dist
        self.dist, self.local_eval = dist, local_eval
        self.R, self.rank, self.world = num_relations, rank, world
        self.draw_sides = draw_sides

    def broadcast_sides(self, side=None):
        import torch
        t = torch.zeros(self.R, dtype=torch.uint8)
        if self.rank == 0:
            s = self.draw_sides() if side is None else side
            t = torch.from_numpy(np.ascontiguousarray(s, np.uint8)).clone()
        if self.world > 1:
            backend = self.dist.get_backend()
            if backend == "nccl":
                t = t.cuda()
            self.dist.broadcast(t, src=0)
        return t.cpu().numpy()

    def computeAt(self, theta, corrupt_side=None):
        import torch
        side = self.broadcast_sides(corrupt_side)
        cost, grad, nact = self.local_eval(theta, side)
        if self.world > 1:
            self.dist.all_reduce(grad)                                    # the one data-path collective
            sc = torch.tensor([cost, float(nact)], dtype=torch.float64, device=grad.device)
            self.dist.all_reduce(sc)
            cost, nact = float(sc[0]), int(round(float(sc[1])))
        return cost, grad, nact


class GradientExchange:
    """The one data-path collective: SUM all-reduce of [gradient | cost, n_active tail] (Pint+4 floats).

    Candidates: NCCL all-reduce, and -- on buffers allocated from torch symmetric memory -- the two-shot
    (reduce-scatter + all-gather over peer loads) and multimem (NVLS: the NVSwitch reduces in the fabric)
    all-reduce kernels.  Measured on 8xB200 for 28.6 MB: NCCL 141 us, two-shot 114 us, multimem 81 us; on
    2 GPUs NCCL wins (74 us).  ``autotune`` times what is available on this job and keeps the fastest."""

    def __init__(self, numel: int, device, n_buffers: int = 2):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.group_name = None
        self.buffers = []
        self.symm = False
        if self.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm
                self.group_name = dist.group.WORLD.group_name
                for _ in range(n_buffers):
                    t = symm.empty(numel, dtype=torch.float32, device=device)
                    symm.rendezvous(t, self.group_name)
                    self.buffers.append(t)
                self.symm = True
            except Exception:
                self.buffers = []
        if not self.buffers:
            self.buffers = [torch.empty(numel, dtype=torch.float32, device=device) for _ in range(n_buffers)]
        self.variant = "nccl"

    def _ops(self):
        d = {"nccl": lambda t: self.dist.all_reduce(t)}
        if self.symm:
            ops = self.torch.ops.symm_mem
            if hasattr(ops, "two_shot_all_reduce_"):
                d["symm_two_shot"] = lambda t: ops.two_shot_all_reduce_(t, "sum", self.group_name)
            if hasattr(ops, "multimem_all_reduce_"):
                d["symm_multimem"] = lambda t: ops.multimem_all_reduce_(t, "sum", self.group_name)
        return d

    def autotune(self, iters=20):
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return {}
        timings = {}
        buf = self.buffers[0]
        for name, op in self._ops().items():
            try:
                buf.zero_()
                for _ in range(3):
                    op(buf)
                torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters):
                    op(buf)
                b.record(); torch.cuda.synchronize()
                ok = 1.0
                ms = a.elapsed_time(b) / iters
            except Exception:
                ok, ms = 0.0, 1e9
            t = torch.tensor([ms, -ok], dtype=torch.float64, device=buf.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)           # slowest rank; any failure disqualifies
            if float(t[1]) == -1.0:
                timings[name] = float(t[0]) * 1e3
        self.variant = min(timings, key=timings.get) if timings else "nccl"
        return timings

    def all_reduce(self, t):
        self._ops()[self.variant](t)

    def bind_fused(self, handles, buffers=None):
        """Hand the symmetric gradient buffers (and their peers' addresses) to the library so that every ntn_eval_dev
        into one of them ends with the full-batch gradient: the exchange runs inside the evaluation, on the library's
        own peer-memory kernels (csrc/exchange.cu), overlapped with the word pull.  Returns False when symmetric memory
        or the entry point is unavailable (the caller then all-reduces after the evaluation)."""
        if not self.symm or self.world == 1:
            return False
        from . import _capi
        lib = _capi.load()
        if not hasattr(lib, "ntn_exchange_bind"):
            return False
        import torch.distributed._symmetric_memory as symm
        torch = self.torch
        buffers = self.buffers if buffers is None else buffers
        ok = True
        # push mode needs a second symmetric allocation per gradient buffer: every rank's scratch receives the other
        # ranks' contributions to the slices it owns (stores only -- no peer loads on the data path)
        want = handles[0].exchange_scratch_numel
        if not hasattr(self, "scratch"):
            self.scratch = {}
        for t in buffers:
            if t.data_ptr() not in self.scratch:
                sc = symm.empty(want, dtype=torch.float32, device=t.device)
                sc.zero_()
                self.scratch[t.data_ptr()] = (sc, symm.rendezvous(sc, self.group_name))
        torch.cuda.synchronize()
        self.dist.barrier()
        for h in handles:
            for t in buffers:
                hdl = symm.rendezvous(t, self.group_name)
                sc, shdl = self.scratch[t.data_ptr()]
                ok = ok and h.exchange_bind(hdl.rank, hdl.world_size, [int(p) for p in hdl.buffer_ptrs],
                                            [int(p) for p in hdl.signal_pad_ptrs], int(getattr(hdl, "multicast_ptr", 0) or 0),
                                            t.numel(), [int(p) for p in shdl.buffer_ptrs], sc.numel())
        # entity-gradient exchange (opt-in, NTN_EXCHANGE_GE=1; default: exchange the word gradient): every handle gets a
        # symmetric entity-gradient buffer + scratch, and the sum over the ranks is taken while the entity pull runs.
        # Measured on 2 x B200 at 20,000 x 10 per GPU: 0.2821 ms against 0.2677 ms -- the exchange kernel that runs
        # beside the entity pull is starved of SM slots (43 us for half of gE) and slows the pull (55 -> 70 us), and the
        # small blocks + tail still need their own exchange after the word pull (18 us).
        import os
        if ok and os.environ.get("NTN_EXCHANGE_GE", "0") != "0":
            if not hasattr(self, "ge"):
                self.ge = []
            for h in handles:
                n = h.exchange_ge_scratch_numel
                ge = symm.empty(n, dtype=torch.float32, device=buffers[0].device)
                gs = symm.empty(n, dtype=torch.float32, device=buffers[0].device)
                ge.zero_(); gs.zero_()
                gh, sh = symm.rendezvous(ge, self.group_name), symm.rendezvous(gs, self.group_name)
                self.ge.append((ge, gs, gh, sh))
                torch.cuda.synchronize()
                self.dist.barrier()
                h.exchange_bind_ge([int(p) for p in gh.buffer_ptrs], [int(p) for p in gh.signal_pad_ptrs],
                                   int(getattr(gh, "multicast_ptr", 0) or 0), n, [int(p) for p in sh.buffer_ptrs], n)
            torch.cuda.synchronize()
            self.dist.barrier()
        return ok


def device_view(ptr: int, n: int, device):
    """torch float32 view of n device floats at ``ptr`` (memory owned by the library)."""
    import torch

    class _Buf:
        pass
    b = _Buf()
    b.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2, "strides": None}
    return torch.as_tensor(b, device=device)


def make_lbfgs_reduce(exchange: GradientExchange, device, stream):
    """reduce(ptr, n) for Handle.lbfgs_minimize.  ``stream`` is the torch stream the handle was bound to with
    Handle.set_stream, so the copies and the collective are ordered after the evaluation that produced the gradient
    and before the optimiser's next kernel.  The library-owned gradient is copied into the exchange buffer (symmetric
    memory when available), all-reduced with the autotuned variant, and copied back."""
    import torch

    def reduce(ptr, n):
        with torch.cuda.stream(stream):
            g = device_view(ptr, n, device)
            buf = exchange.buffers[0]
            buf[:n].copy_(g)
            exchange.all_reduce(buf)
            g.copy_(buf[:n])
    return reduce
