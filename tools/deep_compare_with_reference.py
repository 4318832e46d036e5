"""Deep differential check: the COMPLETE analysis_mem() / analysis_cost() result dictionaries of the reference and of this
repository (every key, every formatted value, the peak-path strings) and the model_arch text (nested module view with
per-module time / FLOP / memory totals) for hand-picked and random configurations.
    python tools/deep_compare_with_reference.py [--cases 0,2,20] [--seed N] [--random K] [--fp8]
"""
import json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from deep_compare_with_reference_lib import SOFT, cw, diff, fz, run   # noqa: E402
import argparse
ap=argparse.ArgumentParser()
ap.add_argument("--cases", default="0,2,4,8,12,15,19", help="indices into compare_with_reference.CASES")
ap.add_argument("--seed", type=int, default=5)
ap.add_argument("--random", type=int, default=12, help="number of random configurations from the fuzzer's sampler")
ap.add_argument("--fp8", action="store_true", help="force fp8 on the random configurations")
args=ap.parse_args()
cases=[cw.CASES[int(i)] for i in args.cases.split(",") if i!=""]
rng=random.Random(args.seed)
for _ in range(args.random):
    c=fz.sample(rng)
    if args.fp8:
        c=list(c[:5])+['a100_pcie']; c[4]=dict(c[4], fp8=True); c=tuple(c)      # a system with fp8 tables
    cases.append(c)
tot=0
for c in cases:
    res=[]
    for impl in ("ref","new"):
        try: res.append(run(impl,c,c[5] if len(c)>5 else "b200_bf16_ceperm"))
        except Exception as e: res.append({"error":type(e).__name__})
    if "error" in res[0] or "error" in res[1]:
        continue
    d=[x for x in diff(res[0],res[1]) if not x[0].endswith('/raw')]; tot+=1      # 'raw' (unformatted numbers) is an addition of this repo
    print("OK " if not d else f"DIFF({len(d)})", c[0], c[1], json.dumps(c[4])[:80], f"| arch lines {len(res[0]['arch'])}")
    for x in d[:6]: print("    ",x)
print("compared",tot,"| values equal as numbers but not as text:",len(SOFT),SOFT[:3])
