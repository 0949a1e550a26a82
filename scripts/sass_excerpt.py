"""Trimmed SASS evidence of the fused encode kernels: python scripts/sass_excerpt.py > profiles/<tag>_sass_excerpt.md

For every rvq_encode_* kernel in librvq_sm100a.so: how often the Blackwell-specific instructions occur (tcgen05 MMA =
UTCHMMA, TMA = UTMALDG, tensor-memory load / store = LDTM / STTM, tcgen05.commit = UTCBAR, bulk copy / reduce =
UBLKCP / UBLKRED, mbarrier = SYNCS) and the instructions around the first MMA group and the first accumulator scan
loop (addresses and encodings stripped)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio_generation_b200", "librvq_sm100a.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(1).strip())
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "UBLKRED", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "FMNMX3", "FMNMX", "RED", "LDGSTS",
        "USETMAXREG", "STL", "LDL"]
print("# SASS excerpt of the fused encode kernels (`cuobjdump -sass audio_generation_b200/librvq_sm100a.so`)\n")
for name, ins in funcs.items():
    if "rvq_encode_" not in name:
        continue
    print(f"## `{demangle(name)}`\n")
    print(f"{len(ins)} instructions.  Occurrences: " + ", ".join(
        f"{k} x{sum(1 for i in ins if re.search(r'(^|\s)' + k + r'(\.|\s|$)', i))}" for k in KEYS) + "\n")
    first = next((i for i, t in enumerate(ins) if "UTCHMMA" in t), None)
    if first is not None:
        last = max(i for i, t in enumerate(ins[first:first + 60]) if "UTCHMMA" in t or "UTCBAR" in t) + first
        print("MMA issue group (one 64-feature slice / ring stage):\n\n```")
        print("\n".join(ins[max(0, first - 6): last + 2]))
        print("```\n")
    scan = next((i for i, t in enumerate(ins) if t.startswith("LDTM.x16")), None)
    if scan is not None:
        print("accumulator scan (two-dimensional running minimum, 32 scores per thread between two LDTM.x16):\n\n```")
        print("\n".join(ins[scan: scan + 64]))
        print("```\n")
