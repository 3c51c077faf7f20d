#!/usr/bin/env python
"""Benchmark of the Sequencer hot path (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5]

A "step" is one full `sequence()` over one synthetic batch of the configured shape (default: BASELINE
config 3, N = 10,000 objects x L = 4,096, ALL_METRICS, scales (1,2,4,8,16)).  Prints ONE JSON line:

  value        whole-job throughput in Gpair/s (one pair = one unordered column pair at one
               (metric, scale) over all L rows; pairs/step = |metrics|*|scales|*N(N-1)/2) with the input
               already resident in HBM; sequence_wall_s is the matching wall time per sequence() call
  e2e          the same metric through the public API `sequence(A)` on a HOST (pinned) array: the H2D copy
               of A and the D2H copy of every result array are inside the timed region
  roofline     the dominant kernel (FP64 DMMA distance tiles): algorithmic FLOPs / measured kernel time,
               against the FP64 tensor peak measured on this pool by tools/fp64_peaks.cu
  cpu_baseline the CPU oracle (a NumPy port of the reference; Julia is not installed) on a bounded sample

The timed region is EXACTLY K steps between two barriers.  If its first step is more than 8 % slower than the
median of its last two (a box that has not settled), the region is timed once more and the line carries both
attempts (`retimed_after_unsettled_first_attempt`).

--impl reference times the oracle port on the host cores instead (rank 0 only).
Multi-GPU (torchrun, one process per GPU): the 20 (metric, scale) groups are sharded over the ranks;
the only exchange is an all-gather of the per-group eta + MST edges before the final fuse.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    1: dict(N=100, L=50, metrics=(0, 1, 2, 3), scales=(1, 2, 4), name="config1: README rand(50,100)"),
    2: dict(N=2000, L=1000, metrics=(1, 3), scales=(1, 2, 4, 8), name="config2: N=2000 x L=1000, EMD+Energy"),
    3: dict(N=10000, L=4096, metrics=(0, 1, 2, 3), scales=(1, 2, 4, 8, 16),
            name="config3: N=10000 objects x L=4096, ALL_METRICS, scales (1,2,4,8,16)"),
    4: dict(N=30000, L=2048, metrics=(1, 3), scales=(1, 2, 4, 8, 16, 32), name="config4: N=30000 x L=2048, EMD+Energy"),
    5: dict(N=20000, L=512, metrics=(0, 2), scales=(1, 2, 4), name="config5: N=20000 x L=512, Euclidean+KL"),
}
FP64_PEAKS = os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")
SEED = 2732          # the seed in the reference's commented-out Random.seed! (test/runtests.jl:10)


def synth(N, L, seed=SEED):
    """rand(Float64, L, N) as the README example (reference README.md:16), object-major (N, L)."""
    return np.random.default_rng(seed).random((N, L))


def pairs_per_step(cfg):
    return len(cfg["metrics"]) * len(cfg["scales"]) * cfg["N"] * (cfg["N"] - 1) // 2


class ClockSampler:
    """SM clock, power and throttle reasons during the timed region (B200_PROFILING.md recipe).

    Read through NVML inside this process (pynvml, the library nvidia-smi itself sits on): a thread polls the four
    values every `period_ms`.  The `nvidia-smi -lms` child process used before stalled this process's kernel launches
    for 50-130 ms at every poll on some boxes -- the slow steps had unchanged kernel times and idle gaps between the
    kernels (profiles/r02_bench_noise.md).  BENCH_SAMPLER=smi selects the child process again; it is also the fallback
    when pynvml is missing.  BENCH_NO_SAMPLER=1 turns sampling off."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0, period_ms=250):
        self.proc = None
        self.nvml = None
        self.lines = []              # smi: csv lines; nvml: (sm, max, watts, [reasons]) tuples
        self.index = index
        self.period_ms = int(os.environ.get("BENCH_SMI_MS", period_ms))
        self.running = False

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x for x in vis.split(",") if x.strip() != ""]
        if ids and all(x.strip().isdigit() for x in ids) and self.index < len(ids):
            return int(ids[self.index])
        return self.index

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return
        if os.environ.get("BENCH_SAMPLER", "nvml") != "smi":
            try:
                import pynvml
                pynvml.nvmlInit()
                self.nvml = pynvml
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
                self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
                self.running = True
                self.t = threading.Thread(target=self._poll, daemon=True)
                self.t.start()
                return
            except Exception:
                self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while self.running:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.lines.append((sm, self.max_sm, pw, [n for n, b in bits.items() if mask & b]))
            except Exception:
                pass
            time.sleep(self.period_ms / 1000.0)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """Start of the timed region: samples from here on are reported (or, if the region is shorter than one
        sampling period, the last sample taken under load during the warm-up)."""
        self.start_idx = len(self.lines)

    def stop(self):
        if self.nvml is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sampler"]}
        if self.nvml is not None:
            self.running = False
            self.t.join(timeout=2)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        lines = self.lines[getattr(self, "start_idx", 0):] or self.lines[-1:]
        for ln in lines:
            if isinstance(ln, tuple):
                sm.append(ln[0]); mx.append(ln[1]); pw.append(ln[2]); reasons.update(ln[3])
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = sorted(sm)[len(sm) // 2:]        # samples under load = upper half
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "period_ms": self.period_ms, "source": "nvml (in-process)" if self.nvml is not None else "nvidia-smi",
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
def cpu_sample(cfg, procs, sample_objects):
    """Oracle (CPU port of the reference) on the first `sample_objects` objects of the same workload."""
    from oracle import seq_oracle as O
    A = synth(cfg["N"], cfg["L"])[:sample_objects].T.copy()         # L x n
    t0 = time.perf_counter()
    O.sequence_oracle_parallel(A, cfg["scales"], cfg["metrics"], procs=procs, faithful_root=False)
    dt = time.perf_counter() - t0
    n = sample_objects
    pairs = len(cfg["metrics"]) * len(cfg["scales"]) * n * (n - 1) // 2
    return pairs / dt / 1e9, dt, pairs


def cpu_faithful(cfg, procs, sizes=(16, 32, 48)):
    """The path the reference really executes for EMD / Energy: one ECDF construction with sorts PER PAIR
    (src/distancemetrics.jl:82-93,143-162,191-223), restated by the oracle's `faithful` mode.  Timed on the first n
    objects for a few n, fitted to t = a + b n^2 (the sizes are smaller than the 100/300/1000
    planned there because the timing has to stay bounded: ~1e3 pairs/s on 8 cores at L = 4096), extrapolated to N."""
    from oracle import seq_oracle as O
    A = synth(cfg["N"], cfg["L"])[:max(sizes)].T.copy()
    pts = []
    for n in sizes:
        t0 = time.perf_counter()
        O.sequence_oracle_parallel(A[:, :n], cfg["scales"], cfg["metrics"], procs=procs, faithful_root=True, faithful=True)
        pts.append((n, time.perf_counter() - t0))
    n2 = np.array([[1.0, n * n] for n, _ in pts])
    a, b = np.linalg.lstsq(n2, np.array([t for _, t in pts]), rcond=None)[0]
    G = len(cfg["metrics"]) * len(cfg["scales"])
    rate = G * 0.5 / max(b, 1e-30) / 1e9                       # pairs per second from the n^2 coefficient, in Gpair/s
    return {"value": rate, "unit": "Gpair/s", "cores": min(procs, G), "kind": "port",
            "sample": f"faithful per-pair ECDF path on the first n = {list(sizes)} objects x L={cfg['L']}, same metrics/scales: "
                      f"{[round(t, 2) for _, t in pts]} s; fit t = {a:.3g} + {b:.3g} n^2",
            "extrapolated_sequence_wall_s": float(a + b * cfg["N"] ** 2),
            "note": "extrapolation by the N^2 law, not a measurement at full size"}


def run_reference(args, cfg):
    """--impl reference: the reference's CPU algorithm (oracle port; Julia absent) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    n = args.sample_objects
    vals = []
    for _ in range(args.warmup):
        cpu_sample(cfg, procs, n)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, dt, pairs = cpu_sample(cfg, procs, n)
        vals.append(v)
    tot = time.perf_counter() - t0
    value = args.steps * pairs / tot / 1e9
    sample = (f"first {n} of {cfg['N']} objects x L={cfg['L']}, same metrics and scales, "
              f"{min(procs, len(cfg['metrics']) * len(cfg['scales']))} processes over the (metric, scale) groups; "
              "pair rate is size-independent to first order (cost ~ N^2 L)")
    out = {"impl": "reference", "metric": METRIC_NAME, "value": value, "unit": "Gpair/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": cfg["name"], "sample": sample},
           "cpu_baseline": {"value": value, "unit": "Gpair/s", "cores": procs, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "Gpair/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "sequence_wall_s_extrapolated": pairs_per_step(cfg) / (value * 1e9)}
    print(json.dumps(out), flush=True)


NCU_TRAFFIC = "profiles/r02_ncu_config3_traffic.json"


def load_ncu_traffic():
    """Per-launch DRAM bytes of the dominant kernels from the committed ncu --set full capture of this command."""
    path = os.path.join(ROOT, NCU_TRAFFIC)
    if not os.path.exists(path):
        return {}
    try:
        return json.load(open(path)).get("kernels", {})
    except Exception:
        return {}


def rank_work(r, N, L, scales, derived):
    """Algorithmic work of the groups / chunks THIS rank computed in one step (from the result's own bookkeeping):
    EMD 2 FP64 instr per pair-element, DMMA tiles 2 flop per pair-element (derived groups run no tiles), MST 8 N^2
    bytes per graph, derive / combine bytes as in DESIGN.md section 4."""
    from sequencerj_b200 import sharding
    npair = N * (N - 1) // 2
    emd_rows = gemm_rows = 0
    graphs = 0
    gemm_groups = 0
    comb = 0.0
    ranges = getattr(r, "chunk_ranges", None) or {}
    for (m, sc), (etas, _) in r.EOSeg.items():
        lens = sharding.chunk_lengths(L, sc)
        a, b = ranges.get((m, sc), (0, len(lens)))
        rows = sum(lens[a:b])
        graphs += (b - a)
        comb += 8.0 * ((b - a) * npair + N * N)
        if m.id in (1, 4, 6):
            emd_rows += rows
        else:
            gemm_rows += rows
            gemm_groups += 1
    graphs += len(r.EOAlgScale)
    if getattr(r, "order", None) is not None:
        graphs += 0            # the final tree is built from the sparse edge union: no N x N read
    gemm_rows -= int(derived) * L                    # a derived group covers all L rows and runs no tiles
    gemm_groups -= int(derived)
    # derive: each derived group reads the upper halves of the next finer scale's chunk matrices and writes its own
    der_bytes = 0.0
    if derived:
        order = sorted(set(scales), key=lambda sc: -sharding.nchunks(L, sc))
        cnt = {}
        for (m, sc) in r.EOSeg:
            if m.id in (0, 2, 3):
                cnt.setdefault(m.id, []).append(sc)
        left = int(derived)
        for mid, scs in cnt.items():
            scs = sorted(scs, key=lambda sc: -sharding.nchunks(L, sc))
            for fine, coarse in zip(scs, scs[1:]):
                if left <= 0:
                    break
                der_bytes += 8.0 * (sharding.nchunks(L, fine) * npair + sharding.nchunks(L, coarse) * N * N)
                left -= 1
    return {"emd_instr": 2.0 * emd_rows * npair, "gemm_flops": 2.0 * max(gemm_rows, 0) * npair, "gemm_groups": max(gemm_groups, 0),
            "mst_bytes": mst_bytes_per_graph(N) * graphs, "graphs": graphs, "derive_bytes": der_bytes, "combine_bytes": comb}


def mst_bytes_per_graph(N):
    """Algorithmic bytes the first list pass of the MST reads per graph (csrc/kernels_mst.cuh).  From N = 6144 (unless
    SEQJ_KNN=rows): the 128 x 128 blocks of the upper triangle incl. the diagonal blocks, once, plus the N/8-entry sample
    of every row; below: every row in full."""
    tiles = N >= 6144 if os.environ.get("SEQJ_KNN") is None else os.environ["SEQJ_KNN"] == "tiles"
    if not tiles:
        return 8.0 * N * N
    nblk = -(-N // 128)
    m = max(64, -(-(N // 8) // 64) * 64)
    return 8.0 * (nblk * (nblk + 1) // 2 * 128 * 128 + N * min(m, N))


METRIC_NAME = "sequence() distance throughput at N=10k,L=4k all metrics (wall-s per call in sequence_wall_s)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--objects", type=int, default=0, help="override N (debugging)")
    ap.add_argument("--length", type=int, default=0, help="override L (debugging)")
    ap.add_argument("--sample-objects", type=int, default=320, help="objects in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-faithful", action="store_true", help="skip the per-pair (faithful) CPU timing")
    ap.add_argument("--sharding", default="group", choices=["group", "chunk"],
                    help="multi-GPU split: whole (metric, scale) groups, or equal-cost chunk ranges + NCCL reduce")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.objects:
        cfg["N"] = args.objects
    if args.length:
        cfg["L"] = args.length
    if args.objects or args.length:
        cfg["name"] += f" [override N={cfg['N']} L={cfg['L']}]"
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch
    import sequencerj_b200 as s
    from sequencerj_b200 import api, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    handle = s.Handle(device=local)
    N, L = cfg["N"], cfg["L"]
    metrics = tuple(s.ALL_METRICS[i] for i in cfg["metrics"])
    scales = cfg["scales"]

    host = torch.empty((N, L), dtype=torch.float64, pin_memory=True)         # object-major = Julia L x N col-major
    host.numpy()[:] = np.maximum(synth(N, L), 2.220446049250313e-16)
    dev = host.cuda(non_blocking=False)
    At = host.numpy()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    device = torch.device("cuda", local)

    def step_dev():
        if world > 1 and args.sharding == "chunk":
            return sharding.sequence_sharded_chunks((N, L), scales, metrics, rank, world, handle, dist, device, dev.data_ptr())
        # group-major: groups packed on the device, ONE NCCL all-gather, fuse from the gathered device buffer
        return sharding.sequence_sharded(At, scales, metrics, rank, world, handle, device_ptr=dev.data_ptr(), dist=dist,
                                         device=device)

    # the nvidia-smi sampler starts BEFORE the warm-up so that its start-up (NVML initialisation takes driver locks)
    # does not land in the timed region; it samples every 250 ms until the timed steps are done (at 100 ms the
    # polling itself cost 3-20 ms per sequence() call of host-side launch stalls, measured with BENCH_SMI_MS /
    # BENCH_NO_SAMPLER)
    # Long steps (>= ~150 ms: config 3 on 1-2 GPUs, config 4) are sampled once a second, short ones every 250 ms: with 250 ms
    # polling about one 230 ms step in 30 ran 10-50 ms longer ON THE DEVICE (profiles/r02_bench_noise.md: 0 of 40 such
    # steps at 1000 ms or without the sampler, 3 of 40 at 250 ms), while a short timed region needs the denser samples
    long_steps = pairs_per_step(cfg) * cfg["L"] / max(world, 1) > 1.5e12
    sampler = ClockSampler(local, (1000 if long_steps else 250) if os.environ.get("BENCH_SAMPLER") == "smi" else 200)
    sampler.start()
    for _ in range(args.warmup):
        r = step_dev()
    step_dev_ms = []

    def timed_region():
        """EXACTLY args.steps steps between two barriers; returns the wall time and the per-step records."""
        barrier()
        sampler.mark()
        t0 = time.perf_counter()
        phases, launches, step_walls, derived, r = {}, 0, [], 0, None
        step_dev_ms.clear()
        for _ in range(args.steps):
            tc = time.perf_counter()
            r = step_dev()
            step_walls.append(round((time.perf_counter() - tc) * 1e3, 1))
            step_dev_ms.append(round(r.timings.get("total", 0.0), 1))   # device-side span of the call (CUDA events)
            if os.environ.get("BENCH_STEP_PHASES"):                      # which phase a slow step was slow in
                print("step phases", {k: round(v, 1) for k, v in r.timings.items() if v and k not in ("input_min", "waves")},
                      file=sys.stderr, flush=True)
            launches += r.launches
            for k, v in r.timings.items():
                if k not in ("input_min", "derived_groups", "waves"):
                    phases[k] = phases.get(k, 0.0) + v
            derived = r.timings.get("derived_groups", 0)
        barrier()
        return time.perf_counter() - t0, phases, launches, step_walls, derived, r

    elapsed, phases, launches, step_walls, derived, r = timed_region()
    # A box that has not settled (seen once, right after another process had released >100 GB of HBM: the first
    # timed steps ran up to 35 % slower than the last ones with identical kernel times, i.e. the GPU sat waiting
    # for the host) is measured again, once, and both attempts are reported.
    first_attempt = None
    unsettled = len(step_walls) >= 3 and step_walls[0] > 1.08 * float(np.median(step_walls[-2:]))
    if dist is not None:
        flag = torch.tensor([1 if unsettled else 0], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        unsettled = bool(flag.item())
    if unsettled:
        first_attempt = {"ms_per_step": elapsed / args.steps * 1e3, "wall_ms_each_step": step_walls}
        elapsed, phases, launches, step_walls, derived, r = timed_region()
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    pairs = pairs_per_step(cfg)
    value = args.steps * pairs / elapsed / 1e9
    order_ref = r.order.copy()

    # ---- end to end through the public API on a host (pinned) array ----
    e2e = None
    if not args.no_e2e:
        A_pub = At.T                         # M x N view, Fortran-ordered like a Julia matrix: no host copy
        def step_e2e():
            if world == 1:
                return s.sequence(A_pub, scales=scales, metrics=metrics, handle=handle)
            if args.sharding == "chunk":
                d = sharding.upload_sharded(At, rank, world, dist, device)       # H2D of A inside the timed region
                torch.cuda.current_stream().synchronize()
                return sharding.sequence_sharded_chunks((N, L), scales, metrics, rank, world, handle, dist, device, d.data_ptr())
            # every rank uploads 1/world of A over its own PCIe link, NVLink all-gather, then as step_dev
            return sharding.sequence_sharded(At, scales, metrics, rank, world, handle, dist=dist, device=device)
        rr = step_e2e()
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        e2e_ph, walls = {}, []
        for _ in range(args.steps):
            tc = time.perf_counter()
            rr = step_e2e()
            walls.append((time.perf_counter() - tc) * 1e3)
            for k in ("total", "h2d", "d2h", "host_upload_ms", "host_local_ms", "host_gather_ms", "host_fuse_ms"):
                if k in rr.timings:
                    e2e_ph[k] = e2e_ph.get(k, 0.0) + rr.timings[k] / args.steps
        barrier()
        el = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        G = len(metrics) * len(scales)
        nch = sum(sharding.nchunks(L, sc) for sc in scales) * len(metrics)
        d2h = nch * (12 + 4 * N) + G * (12 + 4 * N + 16 * N) + (12 + 8 * N + 16 * N) + G * 8 * N
        assert np.array_equal(rr.order, order_ref)
        e2e = {"value": args.steps * pairs / el / 1e9, "unit": "Gpair/s", "h2d_bytes_per_step": 8 * N * L,
               "d2h_bytes_per_step": int(d2h), "sequence_wall_s": el / args.steps,
               "api": ("sequencerj_b200.sequence(A_host) -> seqj_sequence (C ABI), pinned host input" if world == 1 else
                       "sharding.sequence_sharded(A_host): 1/world of A uploaded per rank + NVLink all-gather, seqj_sequence_shard, "
                       "all-gather of the packed group results, seqj_fuse_dev"),
               "device_ms_per_step": {k: round(v, 3) for k, v in e2e_ph.items()}, "wall_ms_each_step": [round(w, 1) for w in walls]}

    # ---- multi-GPU: the sharded result against the unsharded call, and the slowest rank's phases (outside the timed region)
    parity = None
    phases_max = None
    if world > 1:
        keys = list(s._lib.TIMER_NAMES) + ["host_upload_ms", "host_local_ms", "host_gather_ms", "host_fuse_ms", "host_reduce_ms",
                                           "host_finish_ms"]
        t = torch.tensor([phases.get(k, 0.0) / args.steps for k in keys], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases_max = {k: round(float(v), 3) for k, v in zip(keys, t.tolist()) if v > 0}
        etas_all = [None] * world
        dist.all_gather_object(etas_all, {(m.id, sc): e for (m, sc), (e, _) in r.EOAlgScale.items()})
        if rank == 0:
            single = api._sequence(At, scales, metrics, handle, device_ptr=dev.data_ptr())
            merged = {}
            for d_ in etas_all:
                merged.update(d_)
            ref_eta = {(m.id, sc): e for (m, sc), (e, _) in single.EOAlgScale.items()}
            worst = max(abs(merged[k] - ref_eta[k]) / abs(ref_eta[k]) for k in ref_eta) if merged.keys() == ref_eta.keys() else float("inf")
            parity = {"order_identical": bool(np.array_equal(single.order, r.order)), "final_eta_identical": bool(single.eta == r.eta),
                      "groups_compared": len(ref_eta), "worst_group_eta_rel_diff": worst,
                      "ok": bool(np.array_equal(single.order, r.order) and abs(single.eta - r.eta) <= 1e-9 * abs(single.eta)
                                 and worst <= 1e-9)}
        dist.barrier()

    # the same call from a PAGEABLE host array (what a Julia `Matrix` or a plain NumPy array is)
    if e2e is not None and world == 1:
        A_page = np.array(At.T, order="F", copy=True)
        s.sequence(A_page, scales=scales, metrics=metrics, handle=handle)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rp = s.sequence(A_page, scales=scales, metrics=metrics, handle=handle)
        torch.cuda.synchronize()
        elp = time.perf_counter() - t0
        assert np.array_equal(rp.order, order_ref)
        e2e["pageable"] = {"value": args.steps * pairs / elp / 1e9, "unit": "Gpair/s", "sequence_wall_s": elp / args.steps,
                           "h2d_ms": rp.timings.get("h2d", 0.0),
                           "note": "same public call on a pageable (plain NumPy / Julia Matrix) input"}
        del A_page

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- rooflines, computed from the work THIS rank ran (rank 0 prints) ----
    peaks = json.load(open(FP64_PEAKS)) if os.path.exists(FP64_PEAKS) else {}
    hbm = 6533.5
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_src = "fallback 6533.5 GB/s (MEASURED_PEAKS.json absent)"
    if os.path.exists(mp):
        hbm = json.load(open(mp)).get("hbm_gbs", hbm)
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (driver-written)"
    work = rank_work(r, N, L, scales, derived)
    is_cfg3 = args.config == 3 and not (args.objects or args.length)
    ncu = load_ncu_traffic() if (is_cfg3 and world == 1) else {}
    fp64_note = ("builder-measured on this pool by tools/fp64_peaks.cu (profiles/r01_fp64_peaks.json); "
                 "MEASURED_PEAKS.json holds no FP64 figure")
    steps = args.steps

    def roof(kernel, bound, work_per_step, ms_total, peak, unit, scale, peak_source, ncu_key, note=None):
        if ms_total <= 0 or work_per_step <= 0:
            return None
        ach = work_per_step * steps / (ms_total * 1e-3) / scale
        d = {"kernel": kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
             "ms_per_step": ms_total / steps, "work_per_step": work_per_step, "peak_source": peak_source,
             "traffic": None, "traffic_note": "no ncu capture for this workload"}
        if ncu_key in ncu:
            d["traffic"] = ncu[ncu_key]["bytes_per_launch"]
            d["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, STATIC ncu --set full capture of "
                                 f"this command ({NCU_TRAFFIC}); algorithmic bytes per launch {ncu[ncu_key]['algorithmic_bytes']:.4g}")
        if note:
            d["note"] = note
        return d

    gemm_ms = phases.get("dist_l2", 0) + phases.get("dist_kld", 0) + phases.get("dist_energy", 0)
    roofs = {
        "emd": roof("direct_dist_kernel<OP_L1> (EMD tiles: 2 DADD per element)", "fp64-issue", work["emd_instr"],
                    phases.get("dist_emd", 0), peaks.get("dabsdiff_b2_t512_tinstr", 17.6), "Tinstr/s", 1e12, fp64_note, "emd"),
        "gemm": roof("gemm_dist_kernel<L2|KL|Energy> (DMMA tiles + fix-up)", "tensor", work["gemm_flops"], gemm_ms,
                     peaks.get("dmma884_sustained_tflops", 37.06), "TFLOP/s", 1e12, fp64_note, "gemm",
                     note=f"{work['gemm_groups']} tile launches per step, {int(derived)} groups derived instead"),
        "mst": roof("MST stage: knn_sample / knn_tiles / knn_merge kernels + boruvka_knn_kernel (+ rescans, + prim_kernel for graphs "
                    "over the rescan budget)", "hbm", work["mst_bytes"], phases.get("prim", 0), hbm, "GB/s", 1e9, hbm_src, "knn_tiles",
                    note="algorithmic bytes of the list pass: upper-triangle blocks read once (4 N^2 + diagonal blocks) + N^2 for the "
                         "row samples from N = 6144, 8 N^2 below; the stage's time also holds the Boruvka rounds, the rooting and rescans"),
        "derive": roof("derive_kernel<L2|KL|Energy> (coarse scales from the next finer one)", "hbm", work["derive_bytes"],
                       phases.get("derive", 0), hbm, "GB/s", 1e9, hbm_src, "derive",
                       note="8 bytes x (upper halves of the fine matrices read + whole coarse matrices written)"),
        "combine": roof("combine_sym_kernel (eta-weighted accumulation)", "hbm", work["combine_bytes"], phases.get("combine", 0),
                        hbm, "GB/s", 1e9, hbm_src, "combine",
                        note="8 bytes x (upper halves of the chunk matrices read + the group matrix written)"),
    }
    live = {k: v for k, v in roofs.items() if v}
    dominant = max(live, key=lambda k: live[k]["ms_per_step"]) if live else None
    roof_main = dict(live[dominant], dominant_by_time=True) if dominant else None
    if roof_main and roof_main["bound"] == "fp64-issue":
        roof_main["bound_note"] = ("FP64 CUDA-core issue rate (neither HBM nor tensor pipe): L1 between CDFs is not a "
                                   "contraction, the floor is 2 DADD per element")

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        procs = os.cpu_count() or 1
        v, dt, cp = cpu_sample(cfg, procs, args.sample_objects)
        cpu = {"value": v, "unit": "Gpair/s", "cores": procs, "kind": "port",
               "sample": f"vectorised oracle (NumPy port of the reference; Julia absent) on the first {args.sample_objects} of {N} "
                         f"objects x L={L}, same metrics/scales, {dt:.1f} s"}
        if not args.no_faithful:
            cpu["faithful"] = cpu_faithful(cfg, procs)

    out = {"metric": METRIC_NAME, "value": value, "unit": "Gpair/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": cfg["name"], "pairs_per_step": pairs,
                      "cache": "inputs (8*N*L bytes) and every N x N chunk matrix are larger than the 126 MB L2; no flush needed",
                      "parallelism": f"{args.sharding}-major over {world} GPU(s)"},
           "sequence_wall_s": elapsed / args.steps, "wall_ms_each_step": step_walls, "device_ms_each_step": step_dev_ms, "retimed_after_unsettled_first_attempt": first_attempt,
           "e2e": e2e, "gpu_launches": int(launches), "roofline": roof_main,
           "roofline_emd": roofs["emd"], "roofline_gemm": roofs["gemm"], "roofline_mst": roofs["mst"],
           "roofline_derive": roofs["derive"], "roofline_combine": roofs["combine"],
           "roofline_rank": 0, "parity_vs_single_gpu": parity, "phases_ms_per_step_rank_max": phases_max, "dominant_kernel_by_time": dominant, "cpu_baseline": cpu, "clocks": clocks,
           "phases_ms_per_step": {k: v / args.steps for k, v in phases.items()},
           "final_eta": float(r.eta)}
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
