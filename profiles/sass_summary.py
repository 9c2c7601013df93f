"""SASS evidence per kernel: counts of the Blackwell-specific mnemonics in quantize_b200/libqb200.so (cuobjdump -sass).
Usage: python profiles/sass_summary.py > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys

PAT = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)")
KEEP = re.compile(r"^(UTCIMMA|LDTM|UTMALDG|UTMAPF|UTMACCTL|UTCBAR|UTCATOMSWS|UBLKCP|SYNCS|FFMA2|FMUL2|FADD2|IDP|LDGSTS|UTCCP)")
lib = sys.argv[1] if len(sys.argv) > 1 else "quantize_b200/libqb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per = collections.defaultdict(collections.Counter)
tot = collections.Counter()
fn = "?"
for line in out.splitlines():
    if "Function : " in line:
        fn = line.split("Function : ")[1].strip()
        continue
    m = PAT.match(line)
    if m and KEEP.match(m.group(1)):
        per[fn][m.group(1)] += 1
        tot[m.group(1)] += 1
print("# SASS evidence per kernel (cuobjdump -sass quantize_b200/libqb200.so, sm_100a), round 2 — profiles/sass_summary.py")
print("# UTCIMMA = tcgen05.mma kind::i8 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA tensor loads (.IM2COL = im2col mode),")
print("# UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc, SYNCS = mbarrier ops, LDGSTS = cp.async, F*2 = packed fp32, IDP = dp4a")
for k in sorted(tot):
    print("TOTAL", k, tot[k])
def short(f):
    d = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"^void ", "", d)
    d = d.replace("(anonymous namespace)::", "").replace("qb200::", "")
    d = d.split("(")[0]
    return d.replace("false", "0").replace("true", "1").replace(" ", "")
for f in sorted(per):
    print(short(f), " ".join(f"{k}={v}" for k, v in sorted(per[f].items())))
