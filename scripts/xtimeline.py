"""Timeline of one data-parallel evaluation under graph replay (NTN_GRAPH_TIMELINE=1), rank 0's view: kernel groups and
exchange kernels.  torchrun ... scripts/xtimeline.py [config]"""
import os, sys
os.environ["NTN_GRAPH_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
sys.argv = ["bench.py"]
import bench
import torch
import torch.distributed as dist
from neural_tensor_network_for_kb_completion_b200 import _capi
from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory
from neural_tensor_network_for_kb_completion_b200.distributed import GradientExchange, shard_columns
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
bench.SHAPE, bench.BATCH, bench.CONFIG_NAME = bench.CONFIGS[cfg]
kb = bench.build_problem(world, "weak")
gt = bench.global_triples(world, "weak")
tbj = DataFactory.from_kb(kb, gt, bench.CORRUPT, seed=42)
h = _capi.Handle(bench.D, bench.K_SLICES, kb.R, kb.Ne, kb.Nw, gt, bench.LAMBDA, wv_source="theta", device=lr, reg_scale=1.0 / world)
h.set_entity_words(*tbj.entityWordMaps())
h.set_batch(*shard_columns(*tbj.generateNewTrainingBatchJob(), rank, world))
stream = torch.cuda.Stream(device=dev, priority=-1)
torch.cuda.set_stream(stream)
h.set_stream(stream.cuda_stream)
h.set_grad_tail(True)
theta = (np.random.default_rng(0).standard_normal(h.P) * 0.05)
d_t = torch.empty(h.Pint, dtype=torch.float32, device=dev)
h.upload_theta(theta, d_t.data_ptr())
ex = GradientExchange(h.Pint + 4, dev, n_buffers=1)
assert ex.bind_fused([h])
side = np.ones(kb.R, np.uint8)
acc = []
for i in range(30):
    h.eval_dev(d_t.data_ptr(), side, ex.buffers[0].data_ptr(), sync=False)
    torch.cuda.synchronize()
    if i >= 10:
        acc.append(h.debug_fetch("graph_timeline").copy())
t = np.median(np.stack(acc), 0).reshape(-1, 2)
names = [h.lib.ntn_phase_name(i).decode() for i in range(9)] + [f"exchange_{i}" for i in range(len(t) - 9)]
if rank == 0:
    print(cfg, f"world {world}, mode {h.exchange_mode}: graph replay timeline of rank 0 (us, median of 20; begin -> end, duration)")
    for n, (b, e) in sorted(zip(names, t), key=lambda x: x[1][0]):
        if b >= 0:
            print(f"  {n:14s} {b:8.1f} -> {e:8.1f}   {e - b:7.1f}")
dist.destroy_process_group()
