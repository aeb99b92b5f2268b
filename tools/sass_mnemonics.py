"""Counts, per kernel of the built library, the SASS mnemonics that identify the hardware path (profiles/*_sass_mnemonics.txt).
    python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mcan-vqa_b200", "lib", "libmcan_b200.so")
KEEP = re.compile(r"^(UTCHMMA|LDTM|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|SYNCS|HMMA|LDSM|LDGSTS|RED\b|ATOMG|ELECT|MUFU\.EX2)")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = {}
cur = None
counts = collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and cur and KEEP.match(m.group(1)):
        op = m.group(1)
        op = re.sub(r"^(SYNCS)\..*", r"\1", op)
        op = re.sub(r"^(RED|ATOMG)\..*", r"\1", op)
        op = re.sub(r"^(LDSM)\..*", r"\1", op)
        op = re.sub(r"^(LDGSTS)\..*", r"\1", op)
        op = re.sub(r"^(LDTM)\..*", r"\1", op)
        counts[cur][op] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass mcan-vqa_b200/lib/libmcan_b200.so (sm_100a), instruction mnemonics that identify the hardware path,")
print("# counted per kernel (tools/sass_mnemonics.py).  UTCHMMA = tcgen05.mma, UTCHMMA.2CTA = cta_group::2, LDTM = tcgen05.ld (TMEM),")
print("# UTMALDG = TMA tensor load (.2CTA / .MULTICAST variants), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops,")
print("# HMMA.16816 = legacy mma.sync (question-side / guided attention, LSTM recurrence), LDSM = ldmatrix, LDGSTS = cp.async,")
print("# MUFU.EX2 = exp2 (softmax), RED / ATOMG = global atomics.")
print()
rows = []
for mangled, name in zip(counts, dem):
    name = re.sub(r"\(.*", "", name).replace("(anonymous namespace)::", "")
    rows.append((name, "  ".join("%s x%d" % kv for kv in sorted(counts[mangled].items()))))
for name, ops in sorted(rows):
    print("%-70s %s" % (name, ops))
