"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (cuobjdump -sass libiswm_b200.so > sass.txt):
tcgen05 MMA (UTCHMMA / UTCQMMA ...), TMA loads / stores (UTMALDG / UTMASTG), TMEM loads (LDTM), tcgen05 commit (UTCBAR)."""
import collections
import re
import sys

pat = re.compile(r"\b(UTC[A-Z0-9]*MMA[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAREDG[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*|REDG?\.E\.ADD\.F64[.\w]*|RED\.[.\w]*|ATOMG[.\w]*|HMMA[.\w]*)")
cur = None
counts = collections.OrderedDict()
for line in open(sys.argv[1], errors="replace"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for mm in pat.findall(line):
        counts[cur][mm.split(".")[0] + ("." + ".".join(mm.split(".")[1:3]) if "." in mm else "")] += 1
for fn, c in counts.items():
    keys = [k for k in c if k.startswith(("UTC", "UTMA", "LDTM", "STTM"))]
    if not keys:
        continue
    print(fn[:110])
    print("    " + "  ".join(f"{k} x{c[k]}" for k in sorted(c)))
