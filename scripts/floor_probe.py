"""Practical floors on this GPU for the access patterns of the evaluation (torch library kernels, CUDA events, inputs rotated
so that nothing is L2-resident): plain copies of the sizes the streaming kernels move, and row gathers of entity rows."""
import torch, time
dev = "cuda"
def timeit(fn, n=50):
    for _ in range(5): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
NB = 8
for mb in (7, 14, 29, 58, 116):
    n = mb * 1024 * 1024 // 4
    src = [torch.randn(n, device=dev) for _ in range(NB)]
    dst = [torch.empty(n, device=dev) for _ in range(NB)]
    us = timeit(lambda i: dst[i % NB].copy_(src[i % NB]))
    print(f"copy {mb:4d} MB (read+write {2*mb} MB): {us:7.1f} us  -> {2*mb*1.048576/us*1e3:7.0f} GB/s")
Ne, d = 75043, 104
tabs = [torch.randn(Ne, d, device=dev) for _ in range(4)]
for rows in (20000, 75000, 220000):
    idx = [torch.randint(0, Ne, (rows,), device=dev) for _ in range(4)]
    out = [torch.empty(rows, d, device=dev) for _ in range(4)]
    us = timeit(lambda i: torch.index_select(tabs[i % 4], 0, idx[i % 4], out=out[i % 4]))
    print(f"gather {rows:7d} rows x {d*4} B from a {Ne*d*4/1e6:.0f} MB table: {us:7.1f} us -> read {rows*d*4/us/1e3:7.0f} GB/s (+ same written)")
# empty-kernel launch + event overhead
x = torch.zeros(1, device=dev)
print(f"tiny kernel back-to-back: {timeit(lambda i: x.add_(1), 200):.2f} us each")
