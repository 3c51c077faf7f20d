"""Per-launch phase trace of one config-3-shaped sequence() (SEQJ_TRACE=1 prints spans on stderr)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sequencerj_b200 as s
N, L = int(sys.argv[1]), int(sys.argv[2])
A = np.random.default_rng(2732).random((L, N))
os.environ.pop("SEQJ_TRACE", None)
s.sequence(A, scales=(1, 2, 4, 8, 16), metrics=s.ALL_METRICS)
os.environ["SEQJ_TRACE"] = "1"
r = s.sequence(A, scales=(1, 2, 4, 8, 16), metrics=s.ALL_METRICS)
print({k: round(v, 2) for k, v in r.timings.items()})
