"""Multi-GPU check of the in-evaluation gradient exchange (csrc/exchange.cu) against a NCCL all-reduce of the same
per-rank gradients.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1
--master-port 29511 scripts/exchange_check.py [p2p|multimem]   (prints one line per variant on rank 0)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from neural_tensor_network_for_kb_completion_b200 import _capi
    from neural_tensor_network_for_kb_completion_b200.distributed import GradientExchange, shard_columns
    from tests.helpers import perturbed, small_problem, spread

    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    modes = sys.argv[1:] or ["push", "p2p", "multimem"]
    kb, p, batch, side, theta0 = small_problem(Ne=6000, R=6, T=9000, Nw=4100, d=100, batch=6000, corrupt=10, seed=5)
    theta = spread(perturbed(theta0, 0.05), p).astype(np.float32).astype(np.float64)
    stream = torch.cuda.Stream(device=dev, priority=-1)
    torch.cuda.set_stream(stream)
    for mode in modes:
        os.environ["NTN_EXCHANGE"] = mode
        h = _capi.Handle(p.d, p.k, p.R, p.Ne, p.Nw, p.batch_size, p.lam, wv_source="theta", device=lr, reg_scale=1.0 / world)
        h.set_entity_words(p.fwd_off, p.fwd_idx, p.bwd_off, p.bwd_idx, p.len_bwd)
        h.set_batch(*shard_columns(*batch, rank, world))
        h.set_stream(stream.cuda_stream)
        h.set_grad_tail(True)
        Pint = h.Pint
        t_int = torch.empty(Pint, dtype=torch.float32, device=dev)
        h.upload_theta(theta, t_int.data_ptr())
        plain = torch.empty(Pint + 4, dtype=torch.float32, device=dev)
        h.eval_dev(t_int.data_ptr(), side, plain.data_ptr(), sync=False)
        dist.all_reduce(plain)
        torch.cuda.synchronize()
        ex = GradientExchange(Pint + 4, dev, n_buffers=1)
        ok = ex.bind_fused([h])
        if not ok:
            if rank == 0:
                print(f"{mode}: symmetric memory unavailable")
            h.close()
            continue
        buf = ex.buffers[0]
        for _ in range(3):
            buf.fill_(float("nan"))
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            h.eval_dev(t_int.data_ptr(), side, buf.data_ptr(), sync=False)
            torch.cuda.synchronize()
        bad = h.exchange_check()
        diff = float((buf - plain).abs().max())
        scale = float(plain.abs().max())
        # every rank must hold the same bits
        ref = buf.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, buf))
        flags = torch.tensor([float(same), float(bad)], device=dev)
        dist.all_reduce(flags)
        # timing: fused vs evaluation + NCCL
        def timed(fn, n=200):
            for _ in range(5):
                fn()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(n):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / n], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        t_fused = timed(lambda: h.eval_dev(t_int.data_ptr(), side, buf.data_ptr(), sync=False))
        def lib():
            h.eval_dev(t_int.data_ptr(), side, plain.data_ptr(), sync=False)
            dist.all_reduce(plain)
        t_lib = timed(lib)
        t_alone = timed(lambda: h.eval_dev(t_int.data_ptr(), side, plain.data_ptr(), sync=False))
        # the exchange kernel alone: latency (16 bytes) and bandwidth (the whole vector), with and without the end barrier
        probe = {}
        for name, (off, n, last) in {"tiny": (0, 4, 0), "tiny+end": (0, 4, 1), "full": (0, Pint, 0), "full+end": (0, Pint, 1)}.items():
            buf.zero_()
            probe[name] = 1e3 * timed(lambda: h.exchange_probe(buf.data_ptr(), off, n, last, 10), n=20) / 10
        if rank == 0:
            print(f"{mode}: exchange kernel alone, us: " + ", ".join(f"{k} {v:.1f}" for k, v in probe.items()) + f" ({4 * Pint / 1e6:.1f} MB)", flush=True)
        if rank == 0:
            print(f"{mode}: active mode {h.exchange_mode}, world {world}, max|fused - nccl| {diff:.3e} (max |g| {scale:.3e}), "
                  f"identical bits on all ranks: {int(flags[0]) == world}, timeouts: {int(flags[1])}, "
                  f"ms/eval fused {t_fused:.4f}, eval + NCCL {t_lib:.4f}, eval alone {t_alone:.4f}", flush=True)
        h.exchange_unbind()
        h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
