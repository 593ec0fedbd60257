"""Probe (multi-GPU box): latency of the per-evaluation gradient exchange, 7.15 M floats, with NCCL and with torch
symmetric-memory all-reduce variants.  torchrun --nproc-per-node N scripts/allreduce_probe.py"""
import os, sys, time
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 7142580 + 4
dev = torch.device("cuda", lr)


def timeit(fn, iters=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3


x = torch.randn(n, device=dev)
res = {"nccl_allreduce_us": timeit(lambda: dist.all_reduce(x))}
try:
    import torch.distributed._symmetric_memory as symm
    g = dist.group.WORLD
    t = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, g.group_name)
    t.normal_()
    for name in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_"):
        try:
            op = getattr(torch.ops.symm_mem, name)
            res[name + "_us"] = timeit(lambda: op(t, "sum", g.group_name))
        except Exception as e:
            res[name + "_err"] = repr(e)[:200]
    # correctness of the fastest in-place variant
    try:
        t.fill_(float(rank + 1))
        torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", g.group_name)
        torch.cuda.synchronize()
        res["two_shot_ok"] = bool(torch.all(t == world * (world + 1) / 2))
    except Exception as e:
        res["two_shot_check_err"] = repr(e)[:200]
except Exception as e:
    res["symm_err"] = repr(e)[:300]
if rank == 0:
    print("PROBE", world, res, flush=True)
dist.destroy_process_group()
