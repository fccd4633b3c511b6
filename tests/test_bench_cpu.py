"""CPU: host logic of bench.py that no GPU run can exercise deterministically -- the clock sampler's source selection
(in-region NVML samples, else the time-windowed nvidia-smi lines, else one synchronous sample, never an exception) and
the FLOP accounting behind the roofline figures."""
import os
import stat
import sys
import time
import types

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _fake_nvml(fail_polls=False):
    m = types.ModuleType("pynvml")
    m.NVML_CLOCK_SM = 1
    state = {"n": 0, "fail": False}

    def clock(h, t):
        if state["fail"]:
            raise RuntimeError("nvml poll failure")
        state["n"] += 1
        return 1800 + state["n"] % 5
    m.nvmlInit = lambda: None
    m.nvmlDeviceGetHandleByIndex = lambda i: i
    m.nvmlDeviceGetHandleByUUID = lambda u: (_ for _ in ()).throw(RuntimeError("no uuid lookup here"))
    m.nvmlDeviceGetClockInfo = clock
    m.nvmlDeviceGetMaxClockInfo = lambda h, t: 1965
    m.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4
    m._state = state
    m._fail_polls = fail_polls
    return m


@pytest.fixture
def fake_smi(tmp_path, monkeypatch):
    p = tmp_path / "nvidia-smi"
    p.write_text("#!/bin/bash\nwhile true; do echo '0, 1875, 1965, 700.1, 0x4, Not Active, Not Active, Not Active, Active'; sleep 0.05; done\n")
    p.chmod(p.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    return p


def test_clock_sampler_prefers_in_region_nvml_samples(monkeypatch, fake_smi):
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml())
    s = bench.ClockSampler(0)
    time.sleep(0.15)                       # "warm-up": samples before start() must not count
    s.start()
    time.sleep(0.25)
    out = s.stop()
    assert out["source"] == "nvml 20 ms" and 5 <= out["samples"] <= 20
    assert 1800 <= out["sm_mhz"] <= 1804 and out["sm_max_mhz"] == 1965 and out["reasons"] == ["sw_power_cap"]
    assert s.proc.poll() is not None or s.proc.wait(timeout=2) is not None      # the child is gone


def test_clock_sampler_falls_back_to_nvidia_smi_then_to_one_sample(monkeypatch, fake_smi):
    nv = _fake_nvml()
    monkeypatch.setitem(sys.modules, "pynvml", nv)
    s = bench.ClockSampler(0)
    nv._state["fail"] = True               # every in-region NVML poll raises (seen once on a pool box: r02_p_bench_c2.json)
    time.sleep(0.15)
    s.start()
    time.sleep(0.3)
    out = s.stop()
    assert out["source"] == "nvidia-smi -lms 100" and out["samples"] >= 2
    assert out["sm_mhz"] == 1875 and out["reasons"] == ["sw_power_cap"] and out["errors"]
    # no nvidia-smi either: one synchronous NVML sample at the end of the region, and the line says so
    monkeypatch.setenv("PATH", "/nonexistent")
    nv._state["fail"] = False
    s = bench.ClockSampler(0)
    nv._state["fail"] = True
    s.start()
    time.sleep(0.05)
    nv._state["fail"] = False
    out = s.stop()
    assert out["samples"] == 1 and "ONE sample" in out["source"] and out["sm_mhz"] is not None


def test_clock_sampler_without_any_source_reports_instead_of_raising(monkeypatch):
    monkeypatch.setitem(sys.modules, "pynvml", None)      # import pynvml -> ImportError
    monkeypatch.setenv("PATH", "/nonexistent")
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert out["sm_mhz"] is None and out["samples"] == 0 and out["errors"]
    assert bench.ClockSampler(0, enabled=False).stop()["samples"] == 0


def test_flop_accounting_matches_survey():
    # 7.3516 GFLOP per forward of one 197-token sample; 2 forwards x 50 steps = 735.16 GFLOP per motion
    f = bench.flops_per_forward_sample()
    assert abs(f / 1e9 - 7.3516) < 2e-3
    assert abs(2 * 50 * f / 1e9 - 735.16) < 0.2
