"""TEST / BENCH INFRASTRUCTURE ONLY -- recipe for `oracle/_ref/`: the UNMODIFIED reference files of the sampling hot
path, taken from where they lie under /root/reference (build container only) so that the GPU box can time the
reference's own CPU `p_sample_loop` (`bench.py --impl reference`, cpu_baseline.kind = "reference").

  python oracle/build_ref.py            # also run by __graft_entry__.build() when /root/reference is present

What is copied is decided by the reference itself: the harness (oracle/ref_harness.py) imports
`utils.model_util`, `utils.sampler_util`, `diffusion.{gaussian_diffusion,respace}` and `model.mdm` from the reference
tree with its three stubs in place, and every module that import pulled in from /root/reference is copied byte for byte,
same relative path (15 files: diffusion/{gaussian_diffusion,respace,nn,losses}.py, model/mdm.py, utils/{model_util,
sampler_util,misc,loss_util,parser_util}.py and the data_loaders/humanml helpers mdm.py imports).  Nothing is edited;
`oracle/_ref/MANIFEST.json` records the sha256 of every file.  `oracle/_ref/` is git-ignored (reference sources never
enter the history) but travels to the GPU box with the snapshot like the built .so files.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MDM_REFERENCE_SRC", "/root/reference")

_LIST = r'''
import sys, json
sys.path.insert(0, %r)
import os
os.environ["MDM_REFERENCE_ROOT"] = %r
from oracle import ref_harness as rh
rh.load_reference()
root = %r.rstrip("/") + "/"
print("FILES=" + json.dumps(sorted({m.__file__ for m in sys.modules.values()
                                    if getattr(m, "__file__", None) and m.__file__.startswith(root)})))
'''


def build(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "diffusion")):
        return None                      # not the build container: keep whatever oracle/_ref already holds
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", _LIST % (os.path.dirname(HERE), SRC, SRC)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=os.path.dirname(HERE))
    line = [l for l in out.stdout.splitlines() if l.startswith("FILES=")]
    if out.returncode != 0 or not line:
        raise RuntimeError("could not import the reference through oracle/ref_harness.py:\n" + out.stderr[-2000:])
    files = json.loads(line[0][6:])
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for f in files:
        rel = os.path.relpath(f, SRC)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(f, dst)
        manifest[rel] = hashlib.sha256(open(f, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: %d unmodified reference files" % len(files))
    return DST


def available():
    return os.path.isfile(os.path.join(DST, "MANIFEST.json")) and os.path.isdir(os.path.join(DST, "diffusion"))


if __name__ == "__main__":
    print(build(verbose=True))
