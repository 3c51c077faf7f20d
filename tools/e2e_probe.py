"""Five end-to-end sequence() calls on a pinned host array: wall time and the H2D / device totals per call."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s
N, L = 10000, 4096
host = torch.empty((N, L), dtype=torch.float64).pin_memory()
host.copy_(torch.from_numpy(np.random.default_rng(2732).random((N, L))))
A = host.numpy().T
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    print("gpu cpu affinity", [hex(x) for x in pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)], "process", sorted(os.sched_getaffinity(0)))
except Exception as e:
    print("nvml", e)
for i in range(6):
    t = time.perf_counter()
    r = s.sequence(A, scales=(1, 2, 4, 8, 16), metrics=s.ALL_METRICS)
    dt = (time.perf_counter() - t) * 1e3
    print(f"call {i}: wall {dt:7.1f} ms  h2d {r.timings['h2d']:6.2f}  device total {r.timings['total']:7.2f}  d2h {r.timings['d2h']:5.2f}")
