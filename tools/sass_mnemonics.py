"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMA / mbarrier / packed-math use in the shipped library.
    python tools/sass_mnemonics.py [path/to/lib.so] > profiles/sass_mnemonics_<tag>.txt     (CPU only: cuobjdump + c++filt)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "deepsir_b200", "libdeepsir_b200.so")
KEEP = re.compile(r"^(UTCHMMA|LDTM|UTCBAR|UTCATOMSWS|UTMALDG|UBLKCP|SYNCS|FFMA2|FADD2|FMUL2|FMNMX3|MUFU\.EX2|REDUX|LDS\.128|LDL|STL)")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and name and KEEP.match(m.group(1)):
        per[name][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print("SASS mnemonics per kernel, `cuobjdump -sass deepsir_b200/libdeepsir_b200.so` (nvcc 12.9, sm_100a), kernels that use\n"
      "tcgen05 (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc), TMA\n"
      "(UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk), mbarriers (SYNCS.*), packed fp32 math (FFMA2/FADD2/FMUL2),\n"
      "3-input min/max (FMNMX3), warp reductions (REDUX); STL/LDL = local-memory (spill) traffic.\n")
for (k, c), n in zip(per.items(), names):
    if any(re.match(r"^(UTCHMMA|LDTM|UTMALDG|UBLKCP|SYNCS|FFMA2|FADD2|REDUX)", m) for m in c):
        short = re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", ""))
        print(short)
        print("    " + ", ".join(f"{m} x{v}" for m, v in sorted(c.items())))
