#!/usr/bin/env python
"""Bandwidth / latency fit demo on measured collective sweeps (reference: simu_tools/efficency_test/nccl_fit.py, which fits
hard-coded H100 nccl-tests tables).  Here the tables are this box's own sweeps shipped under profiles/: both the in-kernel
NVLink collectives and the NCCL baseline are fitted with the model the cost primitive uses,

    time_us = wire_bytes(op, size, n) / (900 GB/s * efficient_factor) + latency_us * hops ,

and the script prints the fitted parameters, the worst relative residual and the implied bus bandwidth at 256 MiB.

    python tools/nccl_fit.py [profiles/coll_sweep_8gpu.json ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simumax_b200.calib.collectives import fit_bw_latency, wire_bytes  # noqa: E402


def main(paths):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    paths = paths or [os.path.join(root, "profiles", f"coll_sweep_{n}gpu.json") for n in (2, 4, 8)]
    rows = []
    for path in paths:
        d = json.load(open(path))
        n = d["world"]
        for which in ("smx", "nccl"):
            for op, tab in d[which].items():
                fit = fit_bw_latency(op, n, tab["sizes"], tab["times_us"], 900.0)
                big = max(tab["sizes"])
                t_big = tab["times_us"][tab["sizes"].index(big)]
                rows.append((n, which, op, fit["efficient_factor"], fit["latency_us"], fit["max_rel_err"],
                             wire_bytes(op, big, n) / t_big / 1e3))
    print(f"{'gpus':>4} {'impl':>5} {'op':>15} {'eff':>7} {'lat_us':>8} {'max_err':>8} {'busGB/s@max':>12}")
    for r in rows:
        print(f"{r[0]:>4} {r[1]:>5} {r[2]:>15} {r[3]:7.4f} {r[4]:8.3f} {r[5]:8.3f} {r[6]:12.1f}")
    return rows


if __name__ == "__main__":
    main(sys.argv[1:])
