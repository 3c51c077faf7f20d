"""CPU: host-side pieces of bench.py that run on the GPU box without a chance to fail gently -- the in-process NVML
clock sampler (a fake `pynvml` stands in for the driver) and the algorithmic-bytes model of the MST list pass."""
import importlib
import os
import sys
import time
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture
def bench(monkeypatch):
    fake = types.ModuleType("pynvml")
    fake.NVML_CLOCK_SM = 1
    fake.nvmlClocksThrottleReasonHwSlowdown = 0x8
    fake.nvmlClocksThrottleReasonSwPowerCap = 0x4
    fake.nvmlClocksThrottleReasonHwThermalSlowdown = 0x40
    fake.nvmlClocksThrottleReasonSwThermalSlowdown = 0x20
    fake.calls = {"init": 0, "index": None}
    fake.nvmlInit = lambda: fake.calls.__setitem__("init", fake.calls["init"] + 1)
    fake.nvmlDeviceGetHandleByIndex = lambda i: fake.calls.__setitem__("index", i) or "h"
    fake.nvmlDeviceGetMaxClockInfo = lambda h, k: 1965
    fake.nvmlDeviceGetClockInfo = lambda h, k: 1950
    fake.nvmlDeviceGetPowerUsage = lambda h: 800000
    fake.nvmlDeviceGetCurrentClocksThrottleReasons = lambda h: 0x4
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    monkeypatch.delenv("BENCH_SAMPLER", raising=False)
    monkeypatch.delenv("BENCH_NO_SAMPLER", raising=False)
    monkeypatch.delenv("BENCH_SMI_MS", raising=False)
    b = importlib.import_module("bench")
    b._fake_nvml = fake
    return b


def test_nvml_sampler_reports_the_contract_keys(bench, monkeypatch):
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "3,5")
    s = bench.ClockSampler(index=1, period_ms=10)
    s.start()
    time.sleep(0.05)
    s.mark()
    time.sleep(0.08)
    out = s.stop()
    assert bench._fake_nvml.calls["index"] == 5                    # local device 1 of CUDA_VISIBLE_DEVICES=3,5
    assert out["source"].startswith("nvml") and out["samples"] >= 2
    assert out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0 and out["power_w_max"] == 800.0
    assert out["reasons"] == ["sw_power_cap"] and out["period_ms"] == 10


def test_sampler_can_be_switched_off(bench, monkeypatch):
    monkeypatch.setenv("BENCH_NO_SAMPLER", "1")
    s = bench.ClockSampler(index=0)
    s.start()
    s.mark()
    assert s.stop()["reasons"] == ["no sampler"]


def test_mst_bytes_follow_the_list_pass(bench, monkeypatch):
    monkeypatch.delenv("SEQJ_KNN", raising=False)
    N = 10000
    assert bench.mst_bytes_per_graph(4096) == 8.0 * 4096 * 4096            # below 6144: every row in full
    upper = 79 * 80 // 2 * 128 * 128                                        # blocks of the upper triangle incl. the diagonal
    assert bench.mst_bytes_per_graph(N) == 8.0 * (upper + N * 1280)         # + the N/8 sample (rounded up to 64) of every row
    assert bench.mst_bytes_per_graph(N) < 0.66 * 8.0 * N * N
    monkeypatch.setenv("SEQJ_KNN", "rows")
    assert bench.mst_bytes_per_graph(N) == 8.0 * N * N
