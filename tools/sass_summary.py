"""Per-kernel SASS mnemonic evidence (tcgen05 = UTCHMMA/UTCQMMA, TMA = UTMALDG/UTMASTG/UTMAREDG, TMEM = LDTM/STTM,
mbarrier = SYNCS, NVLS = MULTIMEM):   cuobjdump -sass simumax_b200/_kernels_sm100a.so | python tools/sass_summary.py"""
import collections
import re
import sys

SPECIAL = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "SYNCS", "UTCBAR", "MUFU", "MULTIMEM", "RED",
           "LDG", "STG", "HMMA", "FFMA", "UBLKCP", "UBLKRED")
cur, counts = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
print("# cuobjdump -sass simumax_b200/_kernels_sm100a.so (sm_100a): per kernel, instruction total and tensor-core / TMA / TMEM /")
print("# mbarrier / multimem mnemonic counts")
for f, c in counts.items():
    sel = {k: sum(v for kk, v in c.items() if kk.startswith(k)) for k in SPECIAL}
    print(f"{f[:120]}\n    total={sum(c.values())} " + " ".join(f"{k}={v}" for k, v in sel.items() if v))
