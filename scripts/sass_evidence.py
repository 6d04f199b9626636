"""Count the SASS mnemonics that prove which hardware paths each kernel of libmmsa.so uses (tcgen05 MMA = UTC*MMA,
TMEM loads/stores = LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP, mma.sync = HMMA, cp.async = LDGSTS, clusters = UCGABAR).
Runs on the CPU box (cuobjdump only):

    python scripts/sass_evidence.py > profiles/r01_sass_evidence.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200", "mmsa", "libmmsa.so")
PATS = collections.OrderedDict([
    ("UTC*MMA (tcgen05.mma)", re.compile(r"\bUTC\w*MMA")),
    ("LDTM/STTM (tcgen05.ld/st)", re.compile(r"\b(LDTM|STTM)")),
    ("UTMALDG (TMA load)", re.compile(r"\bUTMALDG")),
    ("UTMASTG (TMA store)", re.compile(r"\bUTMASTG")),
    ("UBLKCP (bulk copy)", re.compile(r"\bUBLKCP")),
    ("HMMA (mma.sync)", re.compile(r"\bHMMA")),
    ("LDGSTS (cp.async)", re.compile(r"\bLDGSTS")),
    ("UCGABAR (cluster barrier)", re.compile(r"\bUCGABAR")),
    ("SYNCS (mbarrier)", re.compile(r"\bSYNCS")),
])

out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    for k, p in PATS.items():
        if p.search(line):
            counts[kern][k] += 1

print("# r01: SASS evidence (cuobjdump -sass of libmmsa.so, sm_100a) — instruction counts per kernel\n")
print("Only kernels that use at least one of the listed paths are shown; `scripts/sass_evidence.py` regenerates this file.\n")
print("| kernel | " + " | ".join(PATS) + " |")
print("|---|" + "---:|" * len(PATS))
rows = []
for k, c in counts.items():
    if not any(c.values()):
        continue
    name = demangle(k)
    name = name.replace("(anonymous namespace)::", "").replace("mmsa::", "").replace("void ", "")
    name = re.sub(r"\(.*$", "", name)
    rows.append((name, c))
for name, c in sorted(rows):
    print(f"| `{name[:80]}` | " + " | ".join(str(c.get(p, 0)) for p in PATS) + " |")
