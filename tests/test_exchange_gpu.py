"""The in-evaluation gradient exchange (csrc/exchange.cu) on ONE GPU: G handles play the ranks of a box, their
"symmetric" buffers and signal pads are plain device allocations of the same process (every peer pointer is valid
everywhere), and the G evaluations are enqueued back to back on their own streams, so the exchange kernels of all ranks
are co-resident and synchronise through their flags exactly like G processes over NVLink do.  Checks that every rank ends
with the full-batch gradient, cost and n_active (NTN.java:114-430 is a sum over columns).  The multimem (NVLS) variant
needs a real multicast mapping and is covered by scripts/exchange_check.py under torchrun.

Each case runs in its own process with CUDA_DEVICE_MAX_CONNECTIONS=32: G ranks x 5 streams in ONE process would otherwise
share the default 8 hardware queues, and a rank queued behind another rank's waiting exchange kernel can never arrive (a
false dependency that G separate processes do not have)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# (G, word-pull / entity-pull chunks, variant); "+ge": the sum over the ranks is taken on the entity gradient.  G = 4, 8:
# 4 ranks x 5 streams in one process exhaust the hardware queues -- those run under torchrun (scripts/exchange_check.py)
@pytest.mark.parametrize("G,chunks,mode", [(2, 4, "push"), (2, 1, "push"), (3, 2, "push"), (3, 3, "push"), (2, 4, "p2p"),
                                           (3, 2, "p2p"), (2, 2, "push+ge"), (3, 3, "push+ge"), (2, 1, "p2p+ge")])
def test_loopback_exchange_gives_every_rank_the_full_batch_result(G, chunks, mode):
    # (priority streams draw on a smaller pool of hardware queues: the G ranks of this one-process test give theirs up;
    # eager module loading: a kernel first used by a later rank must not be loaded -- a context-wide synchronisation --
    # while an earlier rank's exchange kernel is waiting for that very rank)
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", NTN_STREAM_PRIORITY="0", CUDA_MODULE_LOADING="EAGER",
               NTN_EXCHANGE_TIMEOUT_MS="30000",      # the ranks of this test capture their graphs one after the other
               NTN_EXCHANGE_CHUNKS=str(chunks),
               NTN_EXCHANGE_BLOCKS="4", NTN_EXCHANGE=mode.split("+")[0], NTN_EXCHANGE_GE_CHUNKS=str(max(chunks, 1)))
    r = subprocess.run([sys.executable, os.path.abspath(__file__), str(G), mode], env=env, cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "loopback ok" in r.stdout


def loopback(G, mode):
    import torch
    from neural_tensor_network_for_kb_completion_b200 import _capi
    from neural_tensor_network_for_kb_completion_b200.distributed import shard_columns
    from tests.helpers import perturbed, small_problem, spread
    from tests.test_gpu_parity import f32, make_handle
    kb, p, batch, side, theta0 = small_problem(Ne=600, R=4, T=900, Nw=420, d=20, batch=800, corrupt=5, seed=21)
    theta = f32(spread(perturbed(theta0, 0.05), p))
    h = make_handle(p, "theta")
    h.set_batch(*batch)
    c_full, g_full, n_full = h.eval(theta, side)
    Pint = h.Pint
    t_int = torch.empty(Pint, dtype=torch.float32, device="cuda")
    h.upload_theta(theta, t_int.data_ptr())
    g_int_full = torch.empty(Pint, dtype=torch.float32, device="cuda")
    h.eval_dev(t_int.data_ptr(), side, g_int_full.data_ptr())
    torch.cuda.synchronize()

    bufs = [torch.full((Pint + 4,), float("nan"), dtype=torch.float32, device="cuda") for _ in range(G)]
    pads = [torch.zeros(9216 // 4, dtype=torch.int32, device="cuda") for _ in range(G)]
    scr = [torch.full((h.exchange_scratch_numel,), float("nan"), dtype=torch.float32, device="cuda") for _ in range(G)]
    ranks = []
    for r in range(G):
        hr = make_handle(p, "theta", reg_scale=1.0 / G)
        hr.set_batch(*shard_columns(*batch, r, G))
        hr.set_grad_tail(True)
        base = mode.split("+")[0]
        hr.exchange_bind(r, G, [b.data_ptr() for b in bufs], [s.data_ptr() for s in pads], 0, Pint + 4,
                         [s.data_ptr() for s in scr] if base == "push" else None, scr[0].numel())
        assert hr.exchange_mode == base
        if mode.endswith("+ge"):
            if r == 0:
                nge = hr.exchange_ge_scratch_numel
                ges = [torch.full((nge,), float("nan"), dtype=torch.float32, device="cuda") for _ in range(G)]
                gscr = [torch.full((nge,), float("nan"), dtype=torch.float32, device="cuda") for _ in range(G)]
                gpads = [torch.zeros(9216 // 4, dtype=torch.int32, device="cuda") for _ in range(G)]
            hr.exchange_bind_ge([b.data_ptr() for b in ges], [s.data_ptr() for s in gpads], 0, nge,
                                [s.data_ptr() for s in gscr] if base == "push" else None, nge)
        ranks.append(hr)
    torch.cuda.synchronize()
    for rep in range(3):                                   # flags return to 0: consecutive evaluations reuse them
        for r in range(G):
            ranks[r].eval_dev(t_int.data_ptr(), side, bufs[r].data_ptr(), sync=False)
        torch.cuda.synchronize()
        for r in range(G):
            assert not ranks[r].exchange_check()
        ref = bufs[0].cpu().numpy()
        for r in range(1, G):
            assert np.array_equal(bufs[r].cpu().numpy(), ref), f"rank {r} differs from rank 0"     # fixed summation order
        cost, nact = _capi.Handle.unpack_tail(ref[Pint:Pint + 4])
        assert nact == n_full
        assert abs(cost - c_full) <= 2e-6 * abs(c_full)
        gi = g_int_full.cpu().numpy()
        assert np.allclose(ref[:Pint], gi, rtol=1e-4, atol=1e-7)
    for hr in ranks:
        hr.close()
    h.close()
    print("loopback ok")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    loopback(int(sys.argv[1]), sys.argv[2])
