"""Per-kernel SASS mnemonic evidence (tcgen05 = UTCHMMA/UTCQMMA, TMA = UTMALDG/UTMASTG/UTMAREDG, TMEM = LDTM/STTM,
mbarrier = SYNCS, NVLS multimem.ld_reduce / multimem.st / multimem.red = LDGMC / STGMC / REDGMC -- "MULTIMEM" is the PTX name and never
appears in SASS):   cuobjdump -sass simumax_b200/_kernels_sm100a.so | python tools/sass_summary.py"""
import collections
import re
import sys

SPECIAL = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "SYNCS", "UTCBAR", "MUFU", "LDGMC", "STGMC", "REDGMC",
           "REDG", "LDG", "STG", "HMMA", "FFMA", "UBLKCP", "UBLKRED")
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
print("# mbarrier / multimem mnemonic counts (multimem.ld_reduce = LDGMC.E.ADD.F32 / LDGMC.E.HPADD.BF16; multimem.st is an ordinary")
print("# STG.E.128.STRONG.SYS on the multicast address: the switch replicates by address, not by opcode)")
for f, c in counts.items():
    # prefix match, except that the plain LDG / STG / REDG counters must not swallow their multicast (..MC) flavours
    sel = {k: sum(v for kk, v in c.items() if kk.startswith(k) and not (k in ("LDG", "STG", "REDG") and kk.split(".")[0].endswith("MC")))
           for k in SPECIAL}
    print(f"{f[:120]}\n    total={sum(c.values())} " + " ".join(f"{k}={v}" for k, v in sel.items() if v))
