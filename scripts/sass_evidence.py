"""cuobjdump -sass libergm_b200.so | python scripts/sass_evidence.py > profiles/r1_sass_evidence.txt
Counts the SASS mnemonics that prove the Blackwell paths per kernel: UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05
ld / st), UTMALDG (TMA tensor loads), UBLKCP (cp.async.bulk), HMMA (legacy mma.sync), RED (fp32 reductions)."""
import collections, re, subprocess, sys
cur = None
cnt = collections.defaultdict(collections.Counter)
pat = re.compile(r"\b(UTC[A-Z]*MMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|HMMA|UTMAPF|UBLKPF|USETMAXREG|REDG?)\b")
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for t in pat.findall(line):
            cnt[cur][t] += 1
keys = list(cnt.keys())
names = subprocess.run(["c++filt"] + keys, capture_output=True, text=True).stdout.split("\n")
print("# cuobjdump -sass ergm_b200/lib/libergm_b200.so: tcgen05 = UTC*MMA + LDTM / STTM (TMEM), TMA = UTMALDG (tensor maps) /")
print("# UBLKCP (bulk copies), legacy tensor path = HMMA (decode slab GEMMs only, HBM-bound at M <= 64)")
seen = set()
for k, n in zip(keys, names):
    short = re.sub(r"\(.*", "", n).replace("void ", "")
    fam = re.sub(r"<.*", "", short)
    c = cnt[k]
    if fam in seen or not any(x in c for x in ("UTCHMMA", "UTMALDG", "UBLKCP", "HMMA", "LDTM")):
        continue
    seen.add(fam)
    print("%-40s %s" % (short[:40], "  ".join("%s=%d" % (a, b) for a, b in sorted(c.items()))))
