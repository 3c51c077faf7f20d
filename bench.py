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
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.lines = []
        self.index = index

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("BENCH_SMI_MS", "250")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """Start of the timed region: samples from here on are reported (or, if the region is shorter than one
        sampling period, the last sample taken under load during the warm-up)."""
        self.start_idx = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[getattr(self, "start_idx", 0):] or self.lines[-1:]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = sorted(sm)[len(sm) // 2:]        # samples under load = upper half
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


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


METRIC_NAME = "sequence() distance throughput at N=10k,L=4k all metrics (wall-s per call in sequence_wall_s)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--objects", type=int, default=0, help="override N (debugging)")
    ap.add_argument("--length", type=int, default=0, help="override L (debugging)")
    ap.add_argument("--sample-objects", type=int, default=320, help="objects in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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

    def allgather(obj):
        outl = [None] * world
        dist.all_gather_object(outl, obj)
        return outl

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    gather_groups = None
    if world > 1:
        def gather_groups(mine, owner):
            return sharding.torch_allgather_groups(mine, owner, metrics, scales, N, rank, world, dist,
                                                   torch.device("cuda", local))

    def step_dev():
        if world > 1 and args.sharding == "chunk":
            return sharding.sequence_sharded_chunks((N, L), scales, metrics, rank, world, handle, dist,
                                                    torch.device("cuda", local), dev.data_ptr())
        return sharding.sequence_sharded(At, scales, metrics, rank, world, handle, allgather, device_ptr=dev.data_ptr(),
                                         gather_groups=gather_groups)

    # the nvidia-smi sampler starts BEFORE the warm-up so that its start-up (NVML initialisation takes driver locks)
    # does not land in the timed region; it samples every 250 ms until the timed steps are done (at 100 ms the
    # polling itself cost 3-20 ms per sequence() call of host-side launch stalls, measured with BENCH_SMI_MS /
    # BENCH_NO_SAMPLER)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        r = step_dev()
    def timed_region():
        """EXACTLY args.steps steps between two barriers; returns the wall time and the per-step records."""
        barrier()
        sampler.mark()
        t0 = time.perf_counter()
        phases, launches, step_walls, derived, r = {}, 0, [], 0, None
        for _ in range(args.steps):
            tc = time.perf_counter()
            r = step_dev()
            step_walls.append(round((time.perf_counter() - tc) * 1e3, 1))
            launches += r.launches
            for k, v in r.timings.items():
                if k not in ("input_min", "derived_groups"):
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
                d = host.to(torch.device("cuda", local), non_blocking=True)      # H2D of A inside the timed region
                return sharding.sequence_sharded_chunks((N, L), scales, metrics, rank, world, handle, dist,
                                                        torch.device("cuda", local), d.data_ptr())
            return sharding.sequence_sharded(At, scales, metrics, rank, world, handle, allgather, gather_groups=gather_groups)
        rr = step_e2e()
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        e2e_ph, walls = {}, []
        for _ in range(args.steps):
            tc = time.perf_counter()
            rr = step_e2e()
            walls.append((time.perf_counter() - tc) * 1e3)
            for k in ("total", "h2d", "d2h"):
                e2e_ph[k] = e2e_ph.get(k, 0.0) + rr.timings.get(k, 0.0) / args.steps
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
               "api": "sequencerj_b200.sequence(A_host) -> seqj_sequence (C ABI), pinned host input",
               "device_ms_per_step": {k: round(v, 3) for k, v in e2e_ph.items()}, "wall_ms_each_step": [round(w, 1) for w in walls]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (FP64 DMMA distance tiles) ----
    peaks = json.load(open(FP64_PEAKS)) if os.path.exists(FP64_PEAKS) else {}
    fp64_peak = peaks.get("dmma884_sustained_tflops", 37.06)
    ngemm = sum(1 for m in cfg["metrics"] if m != 1)
    nemd = sum(1 for m in cfg["metrics"] if m == 1)
    npair1 = N * (N - 1) // 2
    # DMMA flops actually executed: groups whose chunk matrices were derived from the finest scale ran no tiles
    gemm_groups = len(scales) * ngemm / world - derived                             # this rank's share (approx. for N>1)
    gemm_flops = 2.0 * L * npair1 * max(gemm_groups, 0) * args.steps
    gemm_ms = phases.get("dist_l2", 0) + phases.get("dist_kld", 0) + phases.get("dist_energy", 0)
    emd_instr = 2.0 * L * npair1 * len(scales) * nemd * args.steps / world
    emd_ms = phases.get("dist_emd", 0)
    nprim = (sum(sharding.nchunks(L, sc) + 1 for sc in scales) * len(metrics) + 1) * args.steps / world
    prim_bytes = 8.0 * N * N * nprim
    hbm = 6533.5
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        hbm = json.load(open(mp)).get("hbm_gbs", hbm)
    roof = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_config3_metrics.json")
    is_cfg3 = args.config == 3 and not (args.objects or args.length)
    if is_cfg3 and world == 1 and os.path.exists(tpath):
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of the DMMA launches that still run (finest scale:
        # gridDim.y = 16 chunks); the coarser scales are derived and run no tiles
        fin = max(sharding.nchunks(L, sc) for sc in scales)
        ls = [x for x in json.load(open(tpath))["launches"]
              if x["kernel"].startswith("gemm_dist") and x["grid"].replace(" ", "").split(",")[1] == str(fin)]
        if ls:
            traffic = sum(x["dram_read_GB"] + x["dram_write_GB"] for x in ls) * 1e9 / len(ls)
    if gemm_ms > 0:
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12
        roof = {"kernel": "gemm_dist_kernel<L2|KL|Energy> (+ fix-up)", "bound": "tensor", "achieved": ach, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": traffic,
                "traffic_note": ("bytes per launch (finest-scale launches: 16 chunk matrices each), ncu capture in "
                                 "profiles/r01_ncu_config3_metrics.json; algorithmic bytes per launch = 16 x 0.8e9 output "
                                 "+ 0.33e9 operands") if traffic is not None else "no ncu capture for this workload",
                "launches_per_step": int(max(gemm_groups, 0)), "derived_groups_per_step": int(derived),
                "peak_source": "FP64 DMMA.8x8x4 rate measured on this pool (tools/fp64_peaks.cu, profiles/r01_fp64_peaks.json); "
                               "MEASURED_PEAKS.json has no FP64 figure"}
    roof_emd = None
    if emd_ms > 0:
        ach = emd_instr / (emd_ms * 1e-3) / 1e12
        pk = peaks.get("dabsdiff_b2_t512_tinstr", 17.6)
        roof_emd = {"kernel": "direct_dist_kernel<OP_L1> (EMD tiles)", "bound": "fp64-issue", "achieved": ach, "peak": pk, "unit": "Tinstr/s",
                    "frac": ach / pk}
    roof_prim = None
    if phases.get("prim", 0) > 0:
        ach = prim_bytes / (phases["prim"] * 1e-3) / 1e9
        roof_prim = {"kernel": "MST launches of the step: prim_kernel (large batches) + knn_rows_kernel / boruvka_knn_kernel (few graphs)", "bound": "hbm", "achieved": ach,
                     "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                     "note": ("the 124-graph launch alone streams 99.2 GB in 23.4 ms = 4.24 TB/s = 65 % "
                              "(profiles/r01_ncu_config3_metrics.json); the 20 group graphs go through the kNN-filtered "
                              "Boruvka (one bandwidth-bound pass over their matrices)")
                     if is_cfg3 else "all Prim launches of the step (8 N^2 bytes per graph)"}
    if roof is None:
        roof = roof_emd

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        procs = os.cpu_count() or 1
        v, dt, cp = cpu_sample(cfg, procs, args.sample_objects)
        cpu = {"value": v, "unit": "Gpair/s", "cores": procs, "kind": "port",
               "sample": f"oracle (NumPy port of the reference; Julia absent) on the first {args.sample_objects} of {N} objects x "
                         f"L={L}, same metrics/scales, {dt:.1f} s"}

    out = {"metric": METRIC_NAME, "value": value, "unit": "Gpair/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": cfg["name"], "pairs_per_step": pairs,
                      "cache": "inputs (8*N*L bytes) and every N x N chunk matrix are larger than the 126 MB L2; no flush needed",
                      "parallelism": f"{args.sharding}-major over {world} GPU(s)"},
           "sequence_wall_s": elapsed / args.steps, "wall_ms_each_step": step_walls, "retimed_after_unsettled_first_attempt": first_attempt,
           "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
           "roofline_emd": roof_emd, "roofline_prim": roof_prim,
           "dominant_kernel_by_time": max(((phases.get("dist_emd", 0), "direct_dist_kernel<OP_L1> = EMD tiles (FP64 ALU issue-bound, see roofline_emd)"),
                                           (gemm_ms, "gemm_dist_kernel (FP64 tensor pipe, see roofline)"),
                                           (phases.get("prim", 0), "MST kernels (HBM, see roofline_prim)")))[1], "cpu_baseline": cpu, "clocks": clocks,
           "phases_ms_per_step": {k: v / args.steps for k, v in phases.items()},
           "final_eta": float(r.eta)}
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
