"""Exposure factors of the fused compute + collective kernels, from ``tools/fused_check.py`` results.

    exposed = (t_fused - t_gemm) / (t_gemm_then_collective - t_gemm)

i.e. the fraction of the stand-alone (NCCL) collective time that is still visible when the collective runs inside the GEMM
(csrc/gemm_2cta.cu, RS / AG variants).  ``StrategyConfig.tp_overlap_exposed_{ag,rs}`` take these numbers (a float, or a
``{tp_size: value}`` table); ``python -m simumax_b200.calib.fused profiles/fused_ops_*`` prints them.
"""
from __future__ import annotations

import json
import statistics
import sys
from typing import Dict, Iterable


def exposure_from_fused_check(paths: Iterable[str]) -> Dict[int, Dict[str, float]]:
    """{world: {"rs": median exposure of GEMM->reduce-scatter, "ag": ... of all-gather->GEMM, "rs_cases": n, "ag_cases": n}}."""
    out: Dict[int, Dict[str, list]] = {}
    for path in paths:
        d = json.load(open(path))
        acc = out.setdefault(int(d["world"]), {"rs": [], "ag": []})
        for c in d.get("cases", []):
            if c["op"] == "gemm_reduce_scatter" and "fused_tma_ms" in c:
                base, fused, unfused = c["gemm_only_ms"], c["fused_tma_ms"], c["gemm_then_nccl_rs_ms"]
                key = "rs"
            elif c["op"] == "allgather_gemm" and "fused_inkernel_ms" in c:
                base, fused, unfused = c["gemm_only_ms"], c["fused_inkernel_ms"], c["nccl_ag_then_gemm_ms"]
                key = "ag"
            else:
                continue
            if unfused > base:
                acc[key].append(max(0.0, (fused - base) / (unfused - base)))
    return {w: {"rs": statistics.median(v["rs"]) if v["rs"] else None, "ag": statistics.median(v["ag"]) if v["ag"] else None,
                "rs_cases": len(v["rs"]), "ag_cases": len(v["ag"])} for w, v in sorted(out.items())}


def strategy_overrides(table: Dict[int, Dict[str, float]]) -> dict:
    """The two strategy fields as ``{tp_size: exposure}`` tables (worlds without a measurement are left out)."""
    return {"tp_overlap_model": True,
            "tp_overlap_exposed_ag": {str(w): round(v["ag"], 3) for w, v in table.items() if v["ag"] is not None},
            "tp_overlap_exposed_rs": {str(w): round(v["rs"], 3) for w, v in table.items() if v["rs"] is not None}}


if __name__ == "__main__":
    t = exposure_from_fused_check(sys.argv[1:])
    print(json.dumps({"per_world": t, "strategy_overrides": strategy_overrides(t)}, indent=1))
