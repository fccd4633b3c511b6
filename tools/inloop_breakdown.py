"""What each layer kernel costs INSIDE the CUDA-graph sampling loop (programmatic dependent launch, persisting-L2 window) --
which no profiler shows (ncu serialises launches and its replays disturb the L2): time the BASELINE config 2 loop with the
instrumented library (B200MDM_TRACE=1 python -m b200mdm.build) leaving one kernel class out at a time
(B200MDM_DEBUG_SKIP bit mask; results are garbage, only the time matters).   python tools/inloop_breakdown.py"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "motion-diffusion-model_b200", "lib", "libb200mdm_trace.so")
NAMES = {0: "full step", 1: "without fused QKV+attention", 2: "without out-proj+LN GEMM", 4: "without FFN-up GEMM", 8: "without FFN-down+LN GEMM",
         15: "without all four layer kernels"}


def loop_ms(mask):
    env = dict(os.environ, B200MDM_LIB=LIB, B200MDM_DEBUG_SKIP=str(mask))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_loop.py"), "7"], env=env, capture_output=True, text=True).stdout
    m = re.search(r"median ([0-9.]+)", out)
    return float(m.group(1)) if m else float("nan")


if not os.path.exists(LIB):
    sys.path.insert(0, ROOT)
    import importlib
    importlib.import_module("motion-diffusion-model_b200.build").build(trace=True)
base = loop_ms(0)
print("%-40s %8.2f ms per 50-step loop  (%.0f us per step)" % (NAMES[0], base, base * 20))
for mask in (1, 2, 4, 8, 15):
    t = loop_ms(mask)
    n = 8 if mask != 15 else 32
    print("%-40s %8.2f ms  -> that kernel class costs %6.1f us per step = %5.1f us per launch (%4.1f %% of the step)"
          % (NAMES[mask], t, (base - t) * 20, (base - t) * 20 / n, 100 * (base - t) / base))
