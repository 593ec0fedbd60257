"""Begin/end of every kernel group of one evaluation UNDER GRAPH REPLAY (NTN_GRAPH_TIMELINE=1: one-thread stamp kernels
between the groups, also inside the captured graph).  Usage: python scripts/graph_timeline.py [c1|c3|c4]"""
import os, sys
os.environ["NTN_GRAPH_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
sys.argv = ["bench.py"]
import bench
import torch
from neural_tensor_network_for_kb_completion_b200 import _capi
from neural_tensor_network_for_kb_completion_b200.data_factory import DataFactory
bench.SHAPE, bench.BATCH, bench.CONFIG_NAME = bench.CONFIGS[cfg]
kb = bench.build_problem(1)
tbj = DataFactory.from_kb(kb, bench.BATCH, bench.CORRUPT, seed=42)
h = _capi.Handle(bench.D, bench.K_SLICES, kb.R, kb.Ne, kb.Nw, bench.BATCH, bench.LAMBDA, wv_source="theta")
h.set_entity_words(*tbj.entityWordMaps())
h.set_batch(*tbj.generateNewTrainingBatchJob())
stream = torch.cuda.Stream(priority=-1)
h.set_stream(stream.cuda_stream)
theta = (np.random.default_rng(0).standard_normal(h.P) * 0.05)
d_t = torch.empty(h.Pint, dtype=torch.float32, device="cuda"); d_g = torch.empty_like(d_t)
h.upload_theta(theta, d_t.data_ptr())
side = np.ones(kb.R, np.uint8)
acc = []
for i in range(30):
    h.eval_dev(d_t.data_ptr(), side, d_g.data_ptr())
    if i >= 10:
        acc.append(h.debug_fetch("graph_timeline").copy())
t = np.median(np.stack(acc), 0).reshape(-1, 2)
names = [h.lib.ntn_phase_name(i).decode() or f"exchange_{i - 9}" for i in range(len(t))]
print(cfg, "graph replay timeline (us, median of 20 evaluations; begin -> end, duration)")
for n, (b, e) in sorted(zip(names, t), key=lambda x: x[1][0]):
    if b >= 0:
        print(f"  {n:14s} {b:8.1f} -> {e:8.1f}   {e - b:7.1f}")
