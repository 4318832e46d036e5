#!/usr/bin/env python
"""Driver benchmark: PREDICTED vs MEASURED training throughput and peak memory on N B200 GPUs of one node (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                      # own arm  (hand-written sm_100a kernels, runtime/)
    python bench.py --impl reference --gpus N --steps K --warmup W     # reference arm (unmodified reference model + libraries)

What one run does (rank 0 prints ONE JSON line):

* primary case = BASELINE config 2 family (``perf_llama3_8b_tp8``): Llama-3-8B, seq 4096, 8 micro-batches of 1, tensor + sequence
  parallel over all N GPUs (N = 1: single-GPU step), bf16 params, fp32 main grads, AdamW, no recompute.  The global batch is the
  same at every N (strong scaling).  ``value`` = measured whole-job tokens/s (device-timed, max over ranks); ``e2e`` = the same
  through ``train_step`` with pinned-host batches in and a host loss out.
* for every case: what ``PerfLLM`` predicts on the auto-calibrated system config and the resulting ``time_error_pct`` /
  ``peak_mem_error_pct`` (error = (predicted - measured) / measured, the convention of BASELINE.md section 2).
* extra cases that fit N GPUs, each a row (or the <= N-GPU slice of a row) of the reference's published B200 table:
  N = 8: llama3_70b_l12 tp2_dp4 and tp8_dp1, deepseekv2_l4 ep8;  N = 4 / 2: llama3_70b_l12 tp4 / tp2 -- with the published
  Megatron-LM + TE milliseconds next to ours where the row exists.
* N > 1: correctness gates first (in-kernel collectives vs NCCL, TP trainer vs the single-GPU trainer), then the collective
  bandwidth table (in-kernel vs NCCL vs the 900 GB/s NVLink roofline).

Reference arm: the UNMODIFIED reference package (``baseline/_ref``) predicts the same cases through its own ``PerfLLM`` API on its
shipped ``b200_bf16_ceperm.json``; the measured side is ``baseline/lib_trainer.py`` -- the same Megatron-style step on the vendor
libraries the reference calibrates on (cuBLAS, cuDNN attention, NCCL) in stock PyTorch.  Nothing of ``simumax_b200`` is imported
on that path and no ``.so`` of this repo is loaded.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "predicted_vs_measured_train_tokens_per_s_and_peak_mem"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smx", choices=["smx", "reference"])
    ap.add_argument("--cases", default="all", choices=["all", "primary"], help="primary = skip the extra BASELINE cases")
    ap.add_argument("--extra-steps", type=int, default=3, help="timed steps of every extra case (2 warm-up steps)")
    ap.add_argument("--system", default="b200_sm100a_calibrated", help="system config the own arm's prediction uses")
    ap.add_argument("--layers", type=int, default=0, help="debug only: override layer count (marks the run invalid)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-checks", action="store_true", help="skip the multi-GPU correctness gates (debug)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons with nvidia-smi while the timed region runs (B200_PROFILING recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


class Runner:
    """Timing harness shared by both arms: W warm-up steps, K timed steps bracketed by barrier + synchronize, CUDA events on the
    launching stream, max over ranks, a > L2 buffer rewritten before every step."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.a = torch, dist, a
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        assert self.world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run --nproc-per-node {a.gpus}"
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def batches(self, vocab: int, B: int, S: int, n_mb: int, n_sets: int, replica: int):
        """Distinct synthetic token batches per step and micro-batch, in pinned host memory (e2e) and on the device."""
        torch = self.torch
        gen = torch.Generator().manual_seed(1000 + replica)       # the ranks of one model-parallel group share their replica's batch
        host = [[(torch.randint(0, vocab, (B, S), generator=gen).pin_memory(),
                  torch.randint(0, vocab, (B, S), generator=gen).pin_memory()) for _ in range(n_mb)] for _ in range(n_sets)]
        dev = [[(t.cuda().view(-1), y.cuda().view(-1)) for t, y in step] for step in host]
        return host, dev

    def timed(self, fn, steps: int, warmup: int, n_sets: int, sample_clocks: bool = False, launches=None):
        torch = self.torch
        for i in range(warmup):
            fn(i % n_sets)
        sampler = ClockSampler(self.local) if (sample_clocks and self.rank == 0) else None
        self.barrier()
        if sampler:
            sampler.start()
        l0 = launches() if launches else 0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        last = None
        for i in range(steps):
            self.flush.fill_(i & 0xFF)
            last = fn(i % n_sets)
        ev1.record()
        self.barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return {"ms": ms.item() / steps, "wall_ms": wall_ms / steps, "clocks": clocks,
                "launches": (launches() - l0) if launches else None, "last": last}

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([float(v)], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def _err(pred, real):
    return None if pred is None or not real else 100.0 * (pred - real) / real


# ---------------------------------------------------------------------------------------------------------------- own arm
def run_case_smx(R: Runner, case, steps: int, warmup: int, system: str, e2e: bool, sample_clocks: bool):
    import torch
    from simumax_b200.ops import _lib
    from simumax_b200.runtime import cases as C
    mc, dims = C.model_dims(case)
    if R.a.layers:
        dims.layers = case.layers = R.a.layers
    torch.cuda.reset_peak_memory_stats()
    gc.collect(); torch.cuda.empty_cache()
    tr = C.build_trainer(case)
    model_parallel = case.tp * case.pp
    n_sets = max(steps, 1)
    host, dev = R.batches(mc.vocab_size, case.mbs, case.seq, case.mbc, n_sets, replica=R.rank // model_parallel)
    d = R.timed(lambda i: tr.run_microbatches(dev[i]), steps, warmup, n_sets, sample_clocks, launches=_lib.launches)
    loss = float(d["last"].item())
    out = {"case": case.name, "model": case.model, "layers": dims.layers, "seq_len": case.seq, "micro_batch": case.mbs,
           "micro_batches": case.mbc, "parallelism": case.parallelism(), "global_batch": case.mbs * case.mbc * case.dp,
           "trainer": type(tr).__name__, "measured_ms": d["ms"], "tokens_per_s": case.tokens_per_step / d["ms"] * 1e3,
           "gpu_launches": d["launches"], "last_loss": loss, "clocks": d["clocks"]}
    if e2e:
        e = R.timed(lambda i: tr.train_step(host[i]), steps, warmup, n_sets)
        ms = max(e["ms"], e["wall_ms"])
        out["e2e"] = {"value": case.tokens_per_step / ms * 1e3, "unit": "tokens/s", "ms_per_step": ms,
                      "h2d_bytes_per_step": case.mbc * 2 * case.mbs * case.seq * 8, "d2h_bytes_per_step": 4,
                      "api": f"{type(tr).__name__}.train_step(pinned host batches) -> float loss", "last_loss": e["last"]}
    mem = C.device_memory_gib(tr)
    mem = {k: R.max_over_ranks(v) for k, v in mem.items()}
    out["measured_peak_gib"] = mem["peak_gib"]
    out["memory_gib"] = {**{k: round(v, 3) for k, v in mem.items()},
                         "breakdown": {k: round(v / 2 ** 30, 3) for k, v in tr.memory_bytes().items()}}
    C.release(tr)
    del tr, host, dev
    gc.collect(); torch.cuda.empty_cache()
    if R.rank == 0:
        try:
            pred = C.predict(case, system=system, runtime_semantics=True)
            out["predicted"] = pred
            out["time_error_pct"] = _err(pred["iter_ms"], out["measured_ms"])
            out["peak_mem_error_pct"] = _err(pred["peak_gib"], out["measured_peak_gib"])
            ref_sem = C.predict(case, system=system, runtime_semantics=False)
            out["predicted_megatron_semantics"] = {"iter_ms": ref_sem["iter_ms"], "time_error_pct": _err(ref_sem["iter_ms"], out["measured_ms"])}
        except Exception as ex:                                   # never let the model side break the measurement
            out["predicted"] = {"error": f"{type(ex).__name__}: {ex}"}
        pub = C.published(case)
        if pub:
            out["published_megatron"] = {"real_ms": pub[0], "real_gib": pub[1], "reference_pred_ms": pub[2],
                                         "reference_time_error_pct": _err(pub[2], pub[0]),
                                         "speedup_vs_megatron": pub[0] / out["measured_ms"],
                                         "note": "Megatron-LM + TE on 8 x B200 as published by the reference authors (BASELINE.md section 2); quoted, not re-measured"}
    return out


def collectives_table(R: Runner, nbytes: int = 64 << 20):
    """Bus GB/s at ``nbytes`` (ring convention, as nccl-tests): this repo's in-kernel collectives vs NCCL vs 900 GB/s."""
    import torch
    import torch.distributed as dist
    from simumax_b200.calib.collectives import _timed, wire_bytes
    from simumax_b200.parallel.symm import SymmetricWorkspace
    world = R.world
    ws = SymmetricWorkspace(nbytes, nbytes)
    src = torch.empty(nbytes // 2, device="cuda", dtype=torch.bfloat16).normal_()
    dst = torch.empty(nbytes // 2, device="cuda", dtype=torch.bfloat16)
    shard = nbytes // world // 16 * 16
    sh = shard // 2
    nv = bool(ws.has_nvls)
    runs = {
        "all_gather": (lambda: ws.all_gather(shard), lambda: dist.all_gather_into_tensor(dst[:sh * world], src[:sh]), shard * world),
        "reduce_scatter": (lambda: ws.reduce_scatter(shard), lambda: dist.reduce_scatter_tensor(dst[:sh], src[:sh * world]), shard * world),
        "all_reduce": (lambda: ws.all_reduce(nbytes), lambda: dist.all_reduce(src), nbytes),
        "all2all": (lambda: ws.all_to_all(shard), lambda: dist.all_to_all_single(dst[:sh * world], src[:sh * world]), shard * world),
    }
    table = {"bytes": nbytes, "n_gpus": world, "nvls": nv, "variant_policy": "SymmetricWorkspace.pick (same variants as the runtime's exposed collectives)", "nvlink_roofline_gbs": 900.0, "unit": "bus GB/s (bytes*(n-1)/n*scale / time)"}
    for op, (own, nccl, size) in runs.items():
        t_own = _timed(own, 10, dist, torch)
        t_nccl = _timed(nccl, 10, dist, torch)
        wb = wire_bytes(op, size, world)
        table[op] = {"in_kernel_gbs": wb / t_own / 1e3, "nccl_gbs": wb / t_nccl / 1e3, "in_kernel_frac_of_900": wb / t_own / 1e3 / 900.0}
        src.normal_()
    del ws
    return table


def run_smx(a):
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: bench.py measures on a B200 (run it through gpurun)"}))
        return 1
    R = Runner(a)
    from simumax_b200.runtime import cases as C
    from simumax_b200.runtime import selfcheck
    correctness, coll = None, None
    if R.world > 1 and not a.skip_checks:
        correctness = {"collectives_vs_nccl": selfcheck.check_collectives(), "tp_trainer_vs_single_gpu": selfcheck.check_tp_parity()}
        if R.world >= 4:
            correctness["tp2_x_dp_trainer_vs_single_gpu"] = selfcheck.check_tp_parity(tp=2)
        bad = [k for k, v in correctness.items() if not v["ok"]]
        if bad:
            if R.rank == 0:
                print(json.dumps({"error": f"multi-GPU correctness gate failed: {bad}", "correctness": correctness}))
            R.finish()
            return 2
        coll = collectives_table(R)
    elif R.world == 1 and not a.skip_checks:
        correctness = {"recompute_vs_no_recompute": selfcheck.check_recompute_parity()}
    primary = C.primary_case(R.world)
    res = run_case_smx(R, primary, a.steps, a.warmup, a.system, e2e=not a.skip_e2e, sample_clocks=True)
    extras = []
    if a.cases == "all" and not a.layers:
        for case in C.extra_cases(R.world):
            try:
                extras.append(run_case_smx(R, case, a.extra_steps, 2, a.system, e2e=False, sample_clocks=False))
            except Exception as ex:
                extras.append({"case": case.name, "error": f"{type(ex).__name__}: {ex}"[:400]})
                R.barrier()
    if R.rank == 0:
        mc, dims = C.model_dims(primary)
        out = {
            "metric": METRIC, "value": res["tokens_per_s"], "unit": "tokens/s", "n_gpus": R.world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": res["measured_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic random tokens, random-init weights",
            "config": {"model": primary.model, "global_batch": primary.mbs * primary.mbc * primary.dp, "seq_len": primary.seq,
                       "parallelism": primary.parallelism(), "case": primary.name, "micro_batch": primary.mbs, "micro_batches": primary.mbc,
                       "layers": res["layers"], "padded_vocab": dims.vocab, "params": dims.total_numel(),
                       "optimizer": "AdamW fp32 master/m/v, fp32 main grads, grad clip 1.0", "recompute": "none", "attention": "smx fmha fwd+bwd",
                       "baseline_config": "perf_llama3_8b_tp8 (BASELINE.json config 2) at tp = N",
                       "l2": "256 MiB buffer rewritten before every step; a step touches >100 GB of weights/activations"},
            "clocks": res["clocks"], "e2e": res.get("e2e"), "gpu_launches": res["gpu_launches"],
            "time_error_pct": res.get("time_error_pct"), "peak_mem_error_pct": res.get("peak_mem_error_pct"),
            "predicted": res.get("predicted"), "predicted_megatron_semantics": res.get("predicted_megatron_semantics"),
            "measured_peak_gib": res["measured_peak_gib"], "memory_gib": res["memory_gib"], "last_loss": res["last_loss"],
            "cases": [{k: v for k, v in res.items() if k not in ("clocks", "e2e")}] + extras,
            "collectives": coll if coll else {"n_gpus": 1, "note": "single GPU: no NVLink traffic; see MEASURED_PEAKS.json for the HBM copy bandwidth"},
            "correctness": correctness,
        }
        try:     # achieved model FLOP rate against the driver-measured cuBLAS bf16 throughput (MEASURED_PEAKS.json)
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
                                                                                    "note": "fallback: profiling recipe values"}
            mc.layer_num = res["layers"]
            tf = res["tokens_per_s"] * mc.flops_per_token(primary.seq / 2) / 1e12 / R.world     # causal attention: half the S x S scores
            out["roofline"] = {"model_tflops_per_gpu": tf, "cublas_bf16_tflops_burst": peaks.get("bf16_tflops"),
                               "cublas_bf16_tflops_sustained": peaks.get("bf16_tflops_sustained"),
                               "fraction_of_sustained_cublas": tf / peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else None,
                               "fraction_of_nominal_2250": tf / 2250.0, "of": "measured" if os.path.exists(peaks_path) else "fallback"}
        except Exception as e:
            out["roofline"] = {"error": f"{type(e).__name__}: {e}"}
        if a.layers:
            out["invalid"] = "debug run with --layers override"
        out["bench_wall_s"] = round(time.time() - _T0, 1)       # whole process incl. build of the trainer, gates, warm-up, e2e pass
        print(json.dumps(out))
    R.finish()
    return 0


# ---------------------------------------------------------------------------------------------------------------- reference arm
REF_MODELS = {"llama3-8b": "llama3-8b.json", "llama3-70b": "llama3-70b.json"}
REF_CASES = {       # N -> extra cases of the reference arm: (name, model, layers, seq, mbs, mbc, tp, dp, strategy preset)
    8: [("llama3_70b_l12_tp2_pp1_dp4_mbc4", "llama3-70b", 12, 4096, 1, 4, 2, 4, "tp2_pp1_dp4_mbs1"),
        ("llama3_70b_l12_tp8_pp1_dp1_mbc4", "llama3-70b", 12, 4096, 1, 4, 8, 1, "tp8_pp1_dp1_mbs1")],
    4: [("llama3_70b_l12_tp4_pp1_dp1_mbc4", "llama3-70b", 12, 4096, 1, 4, 4, 1, "tp4_pp1_dp2_mbs1")],
    2: [("llama3_70b_l12_tp2_pp1_dp1_mbc4", "llama3-70b", 12, 4096, 1, 4, 2, 1, "tp2_pp1_dp4_mbs1")],
}


def reference_predict(ref_dir, model, layers, seq, mbs, mbc, tp, dp, preset):
    """The unmodified reference's own prediction (its public PerfLLM API, its shipped B200 system config)."""
    import contextlib
    import io
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from simumax.core.config import ModelConfig as RefModel, StrategyConfig as RefStrategy, SystemConfig as RefSystem
    from simumax.core.perf_llm import PerfLLM as RefPerf
    cfg = os.path.join(ref_dir, "configs")
    st = RefStrategy.init_from_config_file(os.path.join(cfg, "strategy", preset + ".json"))
    st.world_size, st.tp_size, st.pp_size = tp * dp, tp, 1
    st.enable_sequence_parallel = tp > 1
    st.seq_len, st.micro_batch_size, st.micro_batch_num = seq, mbs, mbc
    st.enable_recompute = False
    if hasattr(st, "enable_straggler_model"):
        st.enable_straggler_model = False
    if hasattr(st, "__post_init__"):
        st.__post_init__()
    mdl = RefModel.init_from_config_file(os.path.join(cfg, "models", REF_MODELS[model]))
    if layers:
        mdl.layer_num = layers
    mdl.padded_vocab_size, mdl.make_vocab_size_divisible_by = True, 128
    system = RefSystem.init_from_config_file(os.path.join(cfg, "system", "b200_bf16_ceperm.json"))
    p = RefPerf()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        p.configure(strategy_config=st, model_config=mdl, system_config=system)
        p.run_estimate()
        cost = p.analysis_cost()
        mem = p.analysis_mem()
    cd = cost.data if hasattr(cost, "data") else cost
    md = mem.data if hasattr(mem, "data") else mem
    tok_s = cd["throughput_per_accelerator"] * tp * dp
    peak = _parse_peak_gib(md)
    return {"system": "b200_bf16_ceperm (shipped by the reference)", "iter_ms": mbs * mbc * dp * seq / tok_s * 1e3, "tokens_per_s": tok_s,
            "mfu": cd.get("mfu"), "peak_gib": peak}


def _parse_peak_gib(md):
    """Largest ``peak_mem`` string ("12.34 GB") anywhere in an analysis_mem dictionary, in GiB."""
    import re
    units = {"B": 1, "KB": 2 ** 10, "MB": 2 ** 20, "GB": 2 ** 30, "TB": 2 ** 40}
    best = None

    def walk(d):
        nonlocal best
        if isinstance(d, dict):
            for k, v in d.items():
                if k == "peak_mem" and isinstance(v, str):
                    m = re.match(r"\s*([0-9.]+)\s*([KMGT]?B)", v)
                    if m:
                        val = float(m.group(1)) * units[m.group(2)] / 2 ** 30
                        best = val if best is None else max(best, val)
                else:
                    walk(v)
    walk(md)
    return best


def run_case_reference(R: Runner, ref_dir, spec, steps, warmup, e2e, sample_clocks):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    from lib_trainer import Dims, LibLlamaTrainer
    name, model, layers, seq, mbs, mbc, tp, dp, preset = spec
    dims = Dims.from_json(os.path.join(ref_dir, "configs", "models", REF_MODELS[model]), layers=layers or R.a.layers or None)
    if R.a.layers and not layers:
        dims.layers = R.a.layers
    vocab = json.load(open(os.path.join(ref_dir, "configs", "models", REF_MODELS[model])))["vocab_size"]
    tp_group = dp_group = None
    if R.world > 1:
        key = (tp, dp)
        if key not in _REF_GROUPS:
            tps = [dist.new_group(list(range(i * tp, (i + 1) * tp))) for i in range(dp)]
            dps = [dist.new_group(list(range(j, R.world, tp))) for j in range(tp)]
            _REF_GROUPS[key] = (tps, dps)
        tps, dps = _REF_GROUPS[key]
        tp_group = tps[R.rank // tp] if tp > 1 else None
        dp_group = dps[R.rank % tp] if dp > 1 else None
    torch.cuda.reset_peak_memory_stats()
    gc.collect(); torch.cuda.empty_cache()
    tr = LibLlamaTrainer(dims, mbs, seq, tp_group=tp_group, dp_group=dp_group)
    n_sets = max(steps, 1)
    host, dev = R.batches(vocab, mbs, seq, mbc, n_sets, replica=R.rank // tp)
    d = R.timed(lambda i: tr.run_microbatches(dev[i]), steps, warmup, n_sets, sample_clocks)
    tokens = mbs * mbc * dp * seq
    out = {"case": name, "model": model, "layers": dims.layers, "seq_len": seq, "micro_batch": mbs, "micro_batches": mbc,
           "parallelism": f"tp{tp}_pp1_dp{dp}" + ("_sp" if tp > 1 else "") + ("_zero1" if dp > 1 else ""), "global_batch": mbs * mbc * dp,
           "trainer": "baseline/lib_trainer.py::LibLlamaTrainer (torch.matmul / cuDNN SDPA / NCCL)", "measured_ms": d["ms"],
           "tokens_per_s": tokens / d["ms"] * 1e3, "last_loss": float(d["last"].item()), "clocks": d["clocks"]}
    if e2e:
        e = R.timed(lambda i: tr.train_step(host[i]), steps, warmup, n_sets)
        ms = max(e["ms"], e["wall_ms"])
        out["e2e"] = {"value": tokens / ms * 1e3, "unit": "tokens/s", "ms_per_step": ms, "h2d_bytes_per_step": mbc * 2 * mbs * seq * 8,
                      "d2h_bytes_per_step": 4, "api": "LibLlamaTrainer.train_step(pinned host batches) -> float loss", "last_loss": e["last"]}
    out["measured_peak_gib"] = R.max_over_ranks(torch.cuda.max_memory_allocated() / 2 ** 30)
    del tr, host, dev
    gc.collect(); torch.cuda.empty_cache()
    if R.rank == 0:
        try:
            pred = reference_predict(ref_dir, model, layers, seq, mbs, mbc, tp, dp, preset)
            out["predicted"] = pred
            out["time_error_pct"] = _err(pred["iter_ms"], out["measured_ms"])
            out["peak_mem_error_pct"] = _err(pred["peak_gib"], out["measured_peak_gib"])
        except Exception as ex:
            out["predicted"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    return out


_REF_GROUPS: dict = {}


def run_reference(a):
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "simumax", "core")):
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/simumax not installed (run baseline/install_reference.sh; see DESIGN.md)"}))
        return 0
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "reference", "unavailable": "no CUDA device: the reference arm executes its library-kernel step on a GPU"}))
        return 0
    R = Runner(a)
    n = R.world
    primary = (f"llama3_8b_tp{n}_pp1_dp1_mbc8", "llama3-8b", None, 4096, 1, 8, n, 1, "tp8_pp1_dp1_mbs1")
    try:
        res = run_case_reference(R, ref_dir, primary, a.steps, a.warmup, e2e=not a.skip_e2e, sample_clocks=True)
    except Exception as ex:
        if R.rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"reference arm failed: {type(ex).__name__}: {ex}"[:300]}))
        R.finish()
        return 0
    extras = []
    if a.cases == "all" and not a.layers:
        for spec in REF_CASES.get(n, []):
            try:
                extras.append(run_case_reference(R, ref_dir, spec, a.extra_steps, 2, e2e=False, sample_clocks=False))
            except Exception as ex:
                extras.append({"case": spec[0], "error": f"{type(ex).__name__}: {ex}"[:300]})
                R.barrier()
    if R.rank == 0:
        out = {
            "impl": "reference", "metric": METRIC, "value": res["tokens_per_s"], "unit": "tokens/s", "n_gpus": n, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": res["measured_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic random tokens, random-init weights",
            "config": {"model": "llama3-8b", "global_batch": 8, "seq_len": 4096, "parallelism": res["parallelism"], "case": res["case"],
                       "micro_batch": 1, "micro_batches": 8, "layers": res["layers"],
                       "optimizer": "AdamW (torch fused) fp32 master/m/v, fp32 main grads, grad clip 1.0", "recompute": "none",
                       "attention": "F.scaled_dot_product_attention (cuDNN / flash backends)",
                       "l2": "256 MiB buffer rewritten before every step"},
            "reference_arm": "prediction = unmodified reference PerfLLM (baseline/_ref) on its shipped b200_bf16_ceperm.json; measured = the "
                             "same Megatron-style step on the libraries the reference calibrates on (cuBLAS / cuDNN / NCCL) in stock PyTorch "
                             "(baseline/lib_trainer.py) -- Megatron-LM / TransformerEngine are not installed here",
            "clocks": res["clocks"], "e2e": res.get("e2e"), "gpu_launches": 0,
            "time_error_pct": res.get("time_error_pct"), "peak_mem_error_pct": res.get("peak_mem_error_pct"),
            "predicted": res.get("predicted"), "measured_peak_gib": res["measured_peak_gib"], "last_loss": res["last_loss"],
            "cases": [{k: v for k, v in res.items() if k not in ("clocks", "e2e")}] + extras,
        }
        if a.layers:
            out["invalid"] = "debug run with --layers override"
        out["bench_wall_s"] = round(time.time() - _T0, 1)       # whole process incl. build of the trainer, gates, warm-up, e2e pass
        print(json.dumps(out))
    R.finish()
    return 0


_T0 = time.time()


def main():
    a = parse()
    return run_reference(a) if a.impl == "reference" else run_smx(a)


if __name__ == "__main__":
    sys.exit(main() or 0)
