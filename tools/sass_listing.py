"""Per-kernel SASS evidence for profiles/: counts of the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma, LDTM/STTM =
tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UBLKCP = bulk copy, SYNCS = mbarrier, USETMAXREG = setmaxnreg) of the packed fp32
instructions (FFMA2 / FADD2 / FMUL2 = fma/add/mul.f32x2) and of the legacy tensor path (HMMA) in every kernel of libb200mdm.so.   python tools/sass_listing.py > profiles/r02_sass_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "motion-diffusion-model_b200", "lib", "libb200mdm.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "USETMAXREG", "HMMA", "FFMA2", "FADD2", "FMUL2", "MUFU", "STS", "LDS", "STG", "LDG"]
cur, per = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        base = op.split(".")[0]
        if base in MN:
            per[cur][base] += 1
        if op.startswith("UTCHMMA.2CTA"):
            per[cur]["UTCHMMA.2CTA"] += 1
print("%-78s %7s " % ("kernel", "instr") + " ".join("%7s" % m[:7] for m in MN))
for k, c in per.items():
    print("%-78s %7d " % (k[:78], c["_total"]) + " ".join("%7d" % c[m] for m in MN))
