"""How fast can the box's host cores narrow int32 label fields to uint16 (pinned -> pinned),
and how does that compare with the H2D copy time it would save?"""
import os, time, threading
import numpy as np, torch
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
os.system("lscpu | egrep 'Model name|Socket|Core|Thread|NUMA node\\(s\\)' ")
n_fields = 16
src = torch.empty((n_fields, 2048, 2048), dtype=torch.int32).pin_memory()
src.random_(0, 500)
dst = torch.empty((n_fields, 2048, 2048), dtype=torch.int16).pin_memory()
s_np, d_np = src.numpy(), dst.numpy().view(np.uint16)
def work(lo, hi):
    for f in range(lo, hi):
        np.copyto(d_np[f], s_np[f], casting='unsafe')
for T in (1, 2, 4, 8, 16):
    best = 1e9
    for rep in range(3):
        th = [threading.Thread(target=work, args=(i * n_fields // T, (i + 1) * n_fields // T)) for i in range(T)]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        best = min(best, time.perf_counter() - t0)
    print(f"pack threads {T}: {src.numel() * 4 / best / 1e9:.1f} GB/s of int32 read ({best * 1e3 / n_fields:.2f} ms/field)")
dev = torch.empty_like(src, device='cuda'); dev16 = torch.empty_like(dst, device='cuda')
for name, a, b in (("int32", src, dev), ("uint16", dst, dev16)):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        b.copy_(a, non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"H2D {name}: {a.numel() * a.element_size() / dt / 1e9:.1f} GB/s ({dt * 1e3 / n_fields:.2f} ms/field)")
# concurrent: H2D of int32 running while 8 threads pack
for T in (4, 8):
    th = [threading.Thread(target=work, args=(i * n_fields // T, (i + 1) * n_fields // T)) for i in range(T)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dev.copy_(src, non_blocking=True)
    [t.start() for t in th]; [t.join() for t in th]
    t1 = time.perf_counter() - t0
    torch.cuda.synchronize(); t2 = time.perf_counter() - t0
    print(f"concurrent T={T}: pack {t1 * 1e3 / n_fields:.2f} ms/field, H2D int32 {t2 * 1e3 / n_fields:.2f} ms/field")
