"""Host-side pieces of bench.py that can be checked without a GPU: the clock-sampling policy (what is queried inside a
timed region at one and at several GPUs) against a fake NVML."""
import sys
import time
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _fake_nvml(calls):
    m = types.ModuleType("pynvml")
    m.NVML_CLOCK_SM = 1
    m.nvmlClocksThrottleReasonHwSlowdown = 0x8
    m.nvmlClocksThrottleReasonHwThermalSlowdown = 0x40
    m.nvmlClocksThrottleReasonSwThermalSlowdown = 0x20
    m.nvmlClocksThrottleReasonSwPowerCap = 0x4
    m.nvmlInit = lambda: calls.append("init")
    m.nvmlDeviceGetHandleByUUID = lambda uuid: (_ for _ in ()).throw(RuntimeError("no uuid"))
    m.nvmlDeviceGetHandleByIndex = lambda i: ("handle", i)
    m.nvmlDeviceGetMaxClockInfo = lambda h, which: 1965
    m.nvmlDeviceGetClockInfo = lambda h, which: calls.append("clock") or 1950
    m.nvmlDeviceGetCurrentClocksThrottleReasons = lambda h: calls.append("reasons") or 0x4
    return m


def test_clock_sampler_policy(monkeypatch):
    import bench
    calls = []
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml(calls))
    monkeypatch.delenv("RSK_BENCH_SAMPLE_MS", raising=False)

    one = bench.ClockSampler(0, 1)                      # one GPU: clock + reasons inside the region, every 50 ms
    assert one.interval == 0.05 and one.source == "nvml"
    one.sample_adjacent()                               # a no-op with one GPU
    assert calls == ["init"]
    one.start()
    time.sleep(0.12)
    one.pause()
    n = len(one.sm)
    out = one.stop()
    assert n >= 2 and calls.count("clock") == n and calls.count("reasons") == n      # pause() itself takes no sample
    assert out == {"sm_mhz": 1950, "sm_max_mhz": 1965, "reasons": ["sw_power_cap"], "samples": n, "source": "nvml"}

    calls.clear()
    multi = bench.ClockSampler(0, 8)                    # several GPUs: only the clock query inside the region
    assert multi.interval == 1.0
    multi.sample_adjacent()                             # clock + reasons while the warm-up steps execute
    assert calls == ["init", "clock", "reasons"]
    multi.start()
    time.sleep(0.03)
    multi.pause()
    out = multi.stop()
    assert calls.count("reasons") == 1 and calls.count("clock") == 1 + len(multi.sm) and len(multi.sm) >= 1
    assert out["samples"] == len(multi.sm) and out["sm_mhz"] == 1950 and out["reasons"] == ["sw_power_cap"]
    assert out["adjacent"]["sm_mhz"] == [1950]

    calls.clear()
    monkeypatch.setenv("RSK_BENCH_SAMPLE_MS", "off")
    off = bench.ClockSampler(0, 8)
    off.sample_adjacent()
    off.start()
    off.pause()
    assert calls == [] and off.stop()["samples"] == 0
