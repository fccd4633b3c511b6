"""Build recipe for libb200mdm.so (nvcc, sm_100a only, in-tree so the .so travels with the repo snapshot).

    python -m b200mdm.build        # or  __graft_entry__.build()

No GPU is needed to build (nvcc cross-compiles).  The library links the shared CUDA runtime so that it shares
the runtime instance torch has already loaded into the process.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libb200mdm.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "shared",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (looked at %s)" % cand)
    return cand


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200mdm.h")]
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def build(force=False, verbose=False, trace=None):
    """Compile csrc/*.cu into lib/libb200mdm.so.  Returns the library path.
    trace (or B200MDM_TRACE=1): the instrumented variant lib/libb200mdm_trace.so (-DB200_TRACE: clock64 phase stamps for
    tools/trace_*.py; run them with B200MDM_LIB pointing at it)."""
    if trace is None:
        trace = os.environ.get("B200MDM_TRACE", "0") == "1"
    if trace:
        path = os.path.join(LIB_DIR, "libb200mdm_trace.so")
        os.makedirs(LIB_DIR, exist_ok=True)
        proc = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-DB200_TRACE", "-o", path] + sources(), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout)
            raise RuntimeError("nvcc failed (exit %d)" % proc.returncode)
        return path
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB_PATH] + sources()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = proc.stdout
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed (exit %d)" % proc.returncode)
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
