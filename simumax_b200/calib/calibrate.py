"""One-click calibration: measure -> fit -> write ``<name>.json`` that ``PerfLLM`` loads as its system config.

    # single GPU part (operator efficiencies + memory-bound kernels)
    python -m simumax_b200.calib.calibrate ops --out gpurun_out/calib_ops.json
    # multi GPU part, once per group size (collective curves, in-kernel + NCCL baseline)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m simumax_b200.calib.collectives
    # merge everything into a system config (CPU only)
    python -m simumax_b200.calib.calibrate merge --ops gpurun_out/calib_ops.json \
        --coll gpurun_out/coll_sweep_2.json gpurun_out/coll_sweep_4.json gpurun_out/coll_sweep_8.json \
        --out simumax_b200/configs/system/b200_sm100a_calibrated.json

Supplement mode: ``ops --supplement <existing system json>`` measures only the keys that file does not have yet
(the reference's efficiency-cache behaviour, simu_tools/efficency_test/combine_efficiency.py:37-336 and
run_one_click_benchmark.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import Dict, List, Optional

from .collectives import apply_to_system, fit_all, fixed_latency_from_burst
from .shapes import DEFAULT_WORKLOADS, collect_requests
from .system_template import b200_template


def _nominal_tflops(template: dict) -> Dict[str, float]:
    return {op: v["tflops"] for op, v in template["accelerator"]["op"].items()}


def foreign_gpu_processes(device_index: int = 0) -> List[str]:
    """PIDs of other processes holding a context on the GPU we are about to time (calibration on a shared GPU is
    meaningless; reference test_ce_permute_efficiency.py:62-92 refuses to run in that case)."""
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "--query-compute-apps=pid,used_memory", "--format=csv,noheader", "-i", str(device_index)],
                             capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return []
    mine = str(os.getpid())
    return [l.strip() for l in out.splitlines() if l.strip() and l.split(",")[0].strip() != mine]


def cache_path(out_path: str, tag: Optional[str]) -> str:
    """``<out>`` or ``<out stem>_<tag>.json``: the cache namespace of a sweep (reference utils.py:169-174, EFFICIENCY_CACHE_TAG)."""
    if not tag:
        return out_path
    stem, ext = os.path.splitext(out_path)
    return f"{stem}_{tag}{ext or '.json'}"


def plan_requests(param_file: Optional[str] = None, default_sweep: bool = False, workloads=None, log=print):
    """Which (op, shape key) pairs a sweep has to cover: the built-in example workloads, plus the cross product described by a user
    sweep file (``--param-file``, schema of ``shapes.DEFAULT_SWEEP``; reference ``PARAM_FILE``) or the default sweep."""
    from .shapes import collect_requests_tolerant, workloads_from_params
    requests = collect_requests(workloads or DEFAULT_WORKLOADS)
    if param_file or default_sweep:
        params = json.load(open(param_file)) if param_file else None
        extra = collect_requests_tolerant(workloads_from_params(params), log=log)
        for op, table in extra.items():
            requests.setdefault(op, {}).update(table)
    if requests.get("fp8_matmul"):        # the block-scaled recipe (strategy.fp8_recipe = "mxfp8") prices the same linears from its own table
        import dataclasses
        requests["mxfp8_matmul"] = {k: dataclasses.replace(r, op="mxfp8_matmul") for k, r in requests["fp8_matmul"].items()}
    return requests


def run_ops(out_path: str, supplement: Optional[str] = None, limit_per_op: Optional[int] = None,
            workloads=None, log=print, allow_shared_gpu: bool = False, sustained_ms: float = 0.0, param_file: Optional[str] = None,
            default_sweep: bool = False, cache_mode: str = "rebuild", cache_tag: Optional[str] = None,
            library_column: bool = False) -> dict:
    """GPU part: returns {"ops": {op: {key: {...}}}, "bandwidth": {...}, "meta": {...}}.  ``sustained_ms`` > 0 times every op
    at the steady (power-capped) clock instead of the boost clock of a short burst -- see ``measure._time``.

    ``cache_mode`` (reference one_click_common.py:84-94): ``"rebuild"`` measures everything into the cache file (``out_path`` +
    ``cache_tag``); ``"supplement"`` loads that file first and measures only the keys it does not hold yet (an interrupted or
    extended sweep resumes where it stopped).  ``supplement=<system json>`` additionally skips keys that system config already has."""
    import torch
    from . import measure
    assert cache_mode in ("supplement", "rebuild"), f"unsupported cache mode {cache_mode}"
    if sustained_ms > 0:
        measure.SUSTAINED_MS = float(sustained_ms)
    measure.LIBRARY_COLUMN = bool(library_column)
    others = foreign_gpu_processes(torch.cuda.current_device())
    if others and not allow_shared_gpu:
        raise RuntimeError(f"GPU is in use by other processes {others}; calibrate on an idle GPU (or pass --allow-shared-gpu)")
    template = b200_template()
    out_path = cache_path(out_path, cache_tag)
    requests = plan_requests(param_file, default_sweep, workloads, log)
    cached = {}
    if cache_mode == "supplement" and os.path.exists(out_path):
        cached = json.load(open(out_path)).get("ops", {})
        for op in list(requests):
            requests[op] = {k: r for k, r in requests[op].items() if k not in cached.get(op, {})}
        log(f"[calib] cache {out_path}: {sum(len(v) for v in cached.values())} keys reused, {sum(len(v) for v in requests.values())} to measure")
    have = {}
    if supplement and os.path.exists(supplement):
        prev = json.load(open(supplement))
        have = {op: set(v.get("accurate_efficient_factor", {}) or {}) for op, v in prev["accelerator"]["op"].items()}
        for op in list(requests):
            requests[op] = {k: r for k, r in requests[op].items() if k not in have.get(op, set())}
        log(f"[calib] supplement mode: {sum(len(v) for v in requests.values())} keys missing from {supplement}")
    t0 = time.time()
    ops = measure.sweep(requests, _nominal_tflops(template), log=log, limit_per_op=limit_per_op)
    for op, table in cached.items():                  # cache supplement: old keys + the ones measured now
        merged = dict(table)
        merged.update(ops.get(op, {}))
        ops[op] = merged
    vs = [m["vs_cublas"] for m in ops.get("matmul", {}).values() if "vs_cublas" in m]
    if vs:
        vs.sort()
        log(f"[calib] matmul vs cuBLAS on the same operands ({len(vs)} keys): >= 1.0 on {sum(v >= 1.0 for v in vs)}, >= 0.95 on "
            f"{sum(v >= 0.95 for v in vs)}; min {vs[0]:.3f} median {vs[len(vs) // 2]:.3f} max {vs[-1]:.3f}")
    bw = measure.measure_bandwidth_ops()
    for name, rec in bw.items():
        log(f"[calib] bandwidth {name:12s} {rec['gbs']:8.1f} GB/s")
    res = {"ops": ops, "bandwidth": bw, "latency_us": measure.measure_launch_latency_us(),
           "meta": {"cache_mode": cache_mode, "cache_tag": cache_tag, "device": torch.cuda.get_device_name(0), "seconds": time.time() - t0, "sustained_ms": measure.SUSTAINED_MS,
                    "requested": {op: len(v) for op, v in requests.items()}}}
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)
    return res


def merge(ops_path: Optional[str], coll_paths: List[str], out_path: str, base: Optional[str] = None,
          sys_name: str = "b200_sm100a_calibrated", coll_impl: str = "smx", measured_peaks: Optional[str] = None,
          refit_defaults: bool = False) -> dict:
    """CPU part: fold measurements into a system JSON (schema = reference docs/system.md:77-148)."""
    system = json.load(open(base)) if base else b200_template(sys_name)
    system["sys_name"] = sys_name
    acc = system["accelerator"]
    notes = dict(system.get("calibration_notes") or {})
    notes.update({"source": "simumax_b200.calib", "collectives": coll_impl})
    flops_tab = notes.setdefault("flops", {})
    if ops_path:
        res = json.load(open(ops_path))
        for op, table in res["ops"].items():
            cfg = acc["op"].setdefault(op, dict(b200_template()["accelerator"]["op"].get(op) or acc["op"]["default"]))
            tab = cfg.setdefault("accurate_efficient_factor", {}) or {}
            wts = flops_tab.setdefault(op, {})             # flops of every key (kept under calibration_notes): the fallback is flop-weighted
            for key, m in table.items():
                tab[key] = m["efficiency"]
                wts[key] = m["flops"]
            cfg["accurate_efficient_factor"] = tab
            # default for unseen shapes: flop-weighted mean over the MERGED table (keys of the base whose flops are unknown
            # keep the base's fallback: a small supplement must not shift every table miss)
            known = [k for k in tab if k in wts]
            if known and (not base or len(known) == len(tab) or refit_defaults):
                cfg["efficient_factor"] = sum(tab[k] * wts[k] for k in known) / sum(wts[k] for k in known)
        gbps = acc["bandwidth"]["default"]["gbps"]
        lat = res.get("latency_us")                       # measured launch-to-launch floor of a memory-bound kernel, if the sweep has it
        for name, rec in res["bandwidth"].items():
            fresh = name not in acc["bandwidth"]
            entry = acc["bandwidth"].setdefault(name, {"gbps": gbps, "latency_us": 30})
            entry["efficient_factor"] = rec["gbs"] * 1e9 / (entry["gbps"] * 1024 ** 3)
            if lat is not None:
                entry["latency_us"] = float(lat)
            elif not base or fresh:
                entry["latency_us"] = 3.0      # not measured by this sweep: kernel launch + tail of a single-launch elementwise kernel
        means = [acc["op"][o]["efficient_factor"] for o in ("matmul", "sdp_fwd", "sdp_bwd") if o in res["ops"]]
        if means and (not base or refit_defaults):
            acc["op"]["default"]["efficient_factor"] = sum(means) / len(means)
        notes["device"] = res["meta"].get("device")
        notes["sustained_ms"] = res["meta"].get("sustained_ms", 0.0)
    fits, bursts = {}, {}
    for p in coll_paths:
        sw = json.load(open(p))
        fits[int(sw["world"])] = fit_all(sw, nominal_gbps=system["networks"]["high_intra_node"]["bandwidth"]["gbps"], which=coll_impl)
        if sw.get("burst"):
            bursts[int(sw["world"])] = sw["burst"]
    fixed = fixed_latency_from_burst(bursts, fits, which=coll_impl) if bursts else None
    apply_to_system(system, fits, fixed=fixed)
    if measured_peaks and os.path.exists(measured_peaks):
        notes["measured_peaks"] = json.load(open(measured_peaks))
    system["calibration_notes"] = notes
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(system, open(out_path, "w"), indent=1)
    return system


def main(argv=None):
    ap = argparse.ArgumentParser(prog="simumax_b200.calib.calibrate")
    sub = ap.add_subparsers(dest="cmd", required=True)
    o = sub.add_parser("ops", help="single-GPU operator sweeps")
    o.add_argument("--out", default="gpurun_out/calib_ops.json")
    o.add_argument("--supplement", default=None)
    o.add_argument("--limit-per-op", type=int, default=None)
    o.add_argument("--allow-shared-gpu", action="store_true")
    o.add_argument("--param-file", default=None, help="user sweep file (JSON, schema of calib.shapes.DEFAULT_SWEEP): model_list, "
                   "mbs_list, seq_len_list, tp_list, cp_list, ep_list, dtype, long_context")
    o.add_argument("--default-sweep", action="store_true", help="add the built-in zoo sweep (~800 GEMM / 35 attention / 96 grouped keys)")
    o.add_argument("--cache-mode", default="rebuild", choices=["supplement", "rebuild"])
    o.add_argument("--cache-tag", default=None, help="cache namespace: results go to <out stem>_<tag>.json")
    o.add_argument("--library-column", action="store_true", help="also time cuBLAS (torch.matmul) on the same operands for every bf16 GEMM "
                   "key and report the ratio (cublas_ms / cublas_tflops / vs_cublas in the ops file)")
    o.add_argument("--sustained-ms", type=float, default=0.0,
                   help="pre-heat every op for this much device time and time it over windows of a third of it (steady clock)")
    m = sub.add_parser("merge", help="write the system config (no GPU needed)")
    m.add_argument("--ops", default=None)
    m.add_argument("--coll", nargs="*", default=[])
    m.add_argument("--base", default=None)
    m.add_argument("--out", required=True)
    m.add_argument("--name", default="b200_sm100a_calibrated")
    m.add_argument("--coll-impl", default="smx", choices=["smx", "nccl"])
    m.add_argument("--measured-peaks", default="MEASURED_PEAKS.json")
    m.add_argument("--refit-defaults", action="store_true", help="with --base: also recompute the fallback efficiencies from the merged table")
    s = sub.add_parser("shapes", help="list the shapes a sweep needs (CPU)")
    s.add_argument("--param-file", default=None)
    s.add_argument("--default-sweep", action="store_true")
    a = ap.parse_args(argv)
    if a.cmd == "ops":
        run_ops(a.out, a.supplement, a.limit_per_op, allow_shared_gpu=a.allow_shared_gpu, sustained_ms=a.sustained_ms,
                param_file=a.param_file, default_sweep=a.default_sweep, cache_mode=a.cache_mode, cache_tag=a.cache_tag,
                library_column=a.library_column)
    elif a.cmd == "merge":
        sysj = merge(a.ops, a.coll, a.out, a.base, a.name, a.coll_impl, a.measured_peaks, a.refit_defaults)
        print(f"wrote {a.out}: " + ", ".join(f"{op}:{len(v.get('accurate_efficient_factor') or {})}" for op, v in sysj["accelerator"]["op"].items()))
    else:
        for op, t in plan_requests(a.param_file, a.default_sweep, log=lambda *x: None).items():
            print(op, len(t))
            for k in t:
                print("   ", k)
    return 0


if __name__ == "__main__":
    sys.exit(main())
