"""Per-kernel SASS opcode evidence of the shipped library: python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
Counts, per kernel of face_mask_inpaint_b200/libfmi_b200.so, the Blackwell-specific mnemonics that prove the tcgen05 / TMEM /
TMA path (B200_PROFILING.md): UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), UTCQMMA, LDTM / STTM (tcgen05.ld / st), UTCBAR
(tcgen05.commit), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (bulk copy), SYNCS (mbarrier), plus HMMA (legacy mma.sync —
expected 0) and FFMA / MUFU for the SIMT kernels."""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "face_mask_inpaint_b200" / "libfmi_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
ops = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "MUFU",
       "SHFL", "LDG", "STG", "RED", "ATOM"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))[:88]
        if name in per:          # the same template instance compiled in another translation unit: count it once
            cur = collections.Counter()
        else:
            cur = per.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for o in ops:
        if op == o or op.startswith(o + "."):
            if o == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                continue
            cur[o] += 1
            break
print(f"SASS opcode counts per kernel of {lib.name} (sm_100a), cuobjdump -sass; kernels without any listed opcode omitted")
print("%-88s " % "kernel" + " ".join("%8s" % o[:8] for o in ops))
tot = collections.Counter()
for k, c in per.items():
    if not sum(c.values()):
        continue
    tot.update(c)
    print("%-88s " % k + " ".join("%8d" % c[o] for o in ops))
print("%-88s " % "TOTAL" + " ".join("%8d" % tot[o] for o in ops))
