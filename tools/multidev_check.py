"""Single-process multi-GPU (seqj_create with several devices): must equal the single-device result."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s

ndev = int(sys.argv[1]); N, L = int(sys.argv[2]), int(sys.argv[3])
A = np.asfortranarray(np.random.default_rng(5).random((L, N)))
scales = (1, 2, 4, 8, 16) if L >= 1024 else (1, 2, 4)
h1 = s.Handle(0)
hm = s.Handle(list(range(ndev)))
r1 = s.sequence(A, scales=scales, metrics=s.ALL_METRICS, handle=h1)
for rep in range(3):
    t = time.perf_counter(); r1 = s.sequence(A, scales=scales, metrics=s.ALL_METRICS, handle=h1); t1 = time.perf_counter() - t
    t = time.perf_counter(); rm = s.sequence(A, scales=scales, metrics=s.ALL_METRICS, handle=hm); tm = time.perf_counter() - t
same = np.array_equal(r1.order, rm.order) and r1.eta == rm.eta
for key in r1.EOAlgScale:
    same = same and r1.EOAlgScale[key][0] == rm.EOAlgScale[key][0] and np.array_equal(r1.EOAlgScale[key][1].parents, rm.EOAlgScale[key][1].parents)
    same = same and r1.EOSeg[key][0] == rm.EOSeg[key][0]
print(f"{ndev} devices in one process: identical={same}  1 GPU {t1*1e3:.1f} ms  {ndev} GPUs {tm*1e3:.1f} ms  speed-up {t1/tm:.2f}")
print({k: round(v, 1) for k, v in rm.timings.items()})
sys.exit(0 if same else 1)
