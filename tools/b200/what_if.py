"""What a kernel improvement would be worth, priced by the calibrated model: scale an operator efficiency table or a bandwidth entry of the
system config and re-predict the bench cases under runtime semantics.  CPU only.
    python tools/b200/what_if.py                                  # the built-in list (the known gaps of docs/RESULTS.md at library parity)
    python tools/b200/what_if.py --scale sdp_fwd=1.57 --scale sdp_bwd=1.36 --bandwidth norm_bwd=2.5 --gpus 1 8"""
import argparse
import json
import os
import sys
import tempfile
import warnings

warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from simumax_b200.configs import get_simu_system_config  # noqa: E402
from simumax_b200.runtime import cases as C  # noqa: E402

BUILT_IN = [
    ("attention forward at cuDNN parity (x 1.57 at S 4096 GQA)", {"sdp_fwd": 1.57}, {}),
    ("attention backward at cuDNN parity (x 1.36)", {"sdp_bwd": 1.36}, {}),
    ("both attention kernels at cuDNN parity", {"sdp_fwd": 1.57, "sdp_bwd": 1.36}, {}),
    ("every GEMM at same-box burst cuBLAS (x 1.10 on the 8192^3 ratio)", {"matmul": 1.10, "group_matmul": 1.10}, {}),
    ("RMSNorm backward at the elementwise bandwidth (x 2.9)", {}, {"norm_bwd": 2.9}),
    ("fused cross-entropy at the copy bandwidth (x 1.48)", {}, {"ce_fusion": 1.48}),
]


def scaled_config(base_path, ops, bws):
    s = json.load(open(base_path))
    for op, f in ops.items():
        cfg = s["accelerator"]["op"][op]
        cfg["efficient_factor"] *= f
        cfg["accurate_efficient_factor"] = {k: v * f for k, v in (cfg.get("accurate_efficient_factor") or {}).items()}
    for name, f in bws.items():
        entry = s["accelerator"]["bandwidth"].setdefault(name, dict(s["accelerator"]["bandwidth"]["default"]))
        entry["efficient_factor"] *= f
    fd, path = tempfile.mkstemp(suffix=".json")
    with os.fdopen(fd, "w") as f:
        json.dump(s, f)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="b200_sm100a_calibrated")
    ap.add_argument("--scale", action="append", default=[], metavar="op=factor")
    ap.add_argument("--bandwidth", action="append", default=[], metavar="entry=factor")
    ap.add_argument("--gpus", type=int, nargs="*", default=[1, 8])
    a = ap.parse_args()
    base = get_simu_system_config(a.system)
    experiments = BUILT_IN
    if a.scale or a.bandwidth:
        kv = lambda items: {k: float(v) for k, v in (x.split("=") for x in items)}
        experiments = [("custom", kv(a.scale), kv(a.bandwidth))]
    bench_cases = []
    for n in a.gpus:
        bench_cases += [C.primary_case(n)] + [c for c in C.extra_cases(n)]
    ref = {c.name: C.predict(c, system=base)["iter_ms"] for c in bench_cases}
    print("| what if | " + " | ".join(f"{c.name} ({c.world} GPU)" for c in bench_cases) + " |")
    print("|---|" + "---:|" * len(bench_cases))
    print("| calibrated today (ms) | " + " | ".join(f"{ref[c.name]:.1f}" for c in bench_cases) + " |")
    for label, ops, bws in experiments:
        path = scaled_config(base, ops, bws)
        try:
            row = [C.predict(c, system=path)["iter_ms"] for c in bench_cases]
        finally:
            os.unlink(path)
        print(f"| {label} | " + " | ".join(f"{(t / ref[c.name] - 1) * 100:+.1f} %" for t, c in zip(row, bench_cases)) + " |")


if __name__ == "__main__":
    main()
