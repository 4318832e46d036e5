#!/usr/bin/env python
"""Flagship benchmark: measured Llama-3-8B training throughput (tokens/s, whole job) on N B200 GPUs of one node,
reported next to what the analytical model (``PerfLLM.perf()``) predicts for the same system/strategy/model triple.

Own arm  : ``simumax_b200.runtime.trainer.LlamaTrainer`` - every linear / wgrad on the tcgen05 GEMM, tcgen05 flash
           attention, fused norm / SwiGLU / RoPE / CE / AdamW kernels, ZeRO-1 over data-parallel ranks.
Reference: ``--impl reference`` runs the unmodified reference package installed under ``baseline/_ref`` through its
           own ``PerfLLM`` API on its shipped B200 system config and reports ITS predicted tokens/s for this config
           (the reference is an analytical model; it has no training runtime of its own).

Contract (one JSON line on rank 0) is documented in DESIGN.md / the task statement.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASELINE_TOK_S_PER_GPU = 16275.99      # reference llama3-8b tp1_pp2_dp4_mbs1 golden on its B200 config


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smx", choices=["smx", "reference"])
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--seq-len", type=int, default=4096)
    ap.add_argument("--micro-batch", type=int, default=1)
    ap.add_argument("--micro-batches", type=int, default=4, help="gradient-accumulation micro-batches per step per rank")
    ap.add_argument("--layers", type=int, default=0, help="debug only: override layer count (marks the run invalid)")
    ap.add_argument("--attention-bwd", default="auto")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--tp", type=int, default=1, help="tensor(+sequence)-parallel size; >1 runs TPLlamaTrainer on the fused "
                    "all-gather->GEMM / GEMM->reduce-scatter kernels; --gpus / --tp replicas run data parallel (ZeRO-1); default 1 = data parallel")
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


# ----------------------------------------------------------------------------------------------------------------
def predicted(model: str, n_gpus: int, seq: int, mbs: int, mbc: int, tp: int = 1):
    """What our analytical model says about this exact job (dp=N, ZeRO-1, no recompute) on the calibrated config."""
    try:
        from simumax_b200 import PerfLLM
        from simumax_b200.configs import get_simu_model_config, get_simu_system_config
        from simumax_b200.core.config import StrategyConfig
        sys_name = "b200_sm100a_calibrated"
        try:
            sys_path = get_simu_system_config(sys_name)
        except Exception:
            sys_name = "b200_nominal"
            sys_path = get_simu_system_config(sys_name)
        from simumax_b200.configs import get_simu_strategy_config
        st = StrategyConfig.init_from_config_file(get_simu_strategy_config("tp1_pp1_dp8_mbs1"))
        st.world_size, st.tp_size, st.pp_size = n_gpus, tp, 1
        st.enable_sequence_parallel = tp > 1
        st.seq_len, st.micro_batch_size, st.micro_batch_num = seq, mbs, mbc
        st.dp_overlap_model = True      # the runtime overlaps ZeRO-1 collectives with backward / next forward
        dp = n_gpus // tp
        p = PerfLLM()
        p.configure(strategy_config=st, model_config=get_simu_model_config(model), system_config=sys_path)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            p.model_config.padded_vocab_size, p.model_config.make_vocab_size_divisible_by = True, 128
            p.run_estimate()
            r = dict(p.analysis_cost().data)
            mem = p.analysis_mem().data
        r["peak_mem"] = mem.get("peak_mem") if isinstance(mem, dict) else None
        return {"system": sys_name, "tokens_per_s": r["throughput_per_accelerator"] * n_gpus,
                "iter_ms": mbs * mbc * dp * seq / (r["throughput_per_accelerator"] * n_gpus) * 1e3,
                "mfu": r.get("mfu"), "peak_mem": r.get("peak_mem")}
    except Exception as e:                                   # never let the model side break the measurement
        return {"error": f"{type(e).__name__}: {e}"}


def run_reference(a):
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "simumax")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/simumax not installed (see DESIGN.md)"}))
        return
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, ref_dir)
    t0 = time.perf_counter()
    try:
        from simumax.core.perf_llm import PerfLLM as RefPerf                          # reference public API
        from simumax.core.config import StrategyConfig as RefStrategy, SystemConfig as RefSystem, ModelConfig as RefModel
        cfg = os.path.join(ref_dir, "configs")
        if not os.path.isdir(cfg):
            cfg = "/root/reference/configs"
        st_path = os.path.join(cfg, "strategy", "tp1_pp1_dp8_mbs1.json")
        strategy = RefStrategy.init_from_config_file(st_path)
        strategy.world_size, strategy.tp_size, strategy.pp_size = a.gpus, 1, 1
        strategy.seq_len, strategy.micro_batch_size, strategy.micro_batch_num = a.seq_len, a.micro_batch, a.micro_batches
        strategy.__post_init__() if hasattr(strategy, "__post_init__") else None
        model = RefModel.init_from_config_file(os.path.join(cfg, "models", f"{a.model}.json"))
        model.padded_vocab_size, model.make_vocab_size_divisible_by = True, 128
        system = RefSystem.init_from_config_file(os.path.join(cfg, "system", "b200_bf16_ceperm.json"))
        p = RefPerf()
        p.configure(strategy_config=strategy, model_config=model, system_config=system)
        p.run_estimate()
        cost = p.analysis_cost()
        d = cost.data if hasattr(cost, "data") else cost
        tok = d["throughput_per_accelerator"] * a.gpus
        ms = a.micro_batch * a.micro_batches * a.gpus * a.seq_len / tok * 1e3
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": f"reference run failed: {type(e).__name__}: {e}"[:300]}))
        return
    wall = time.perf_counter() - t0
    print(json.dumps({
        "impl": "reference", "metric": "llama3_8b_train_tokens_per_s", "value": tok, "unit": "tokens/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": tok / (BASELINE_TOK_S_PER_GPU * a.gpus), "dtype": "bf16",
        "data": "analytical prediction by the unmodified reference PerfLLM on its shipped b200_bf16_ceperm system config "
                "(the reference has no training runtime; nothing is executed on the GPU)",
        "config": {"model": a.model, "global_batch": a.micro_batch * a.micro_batches * a.gpus, "seq_len": a.seq_len,
                   "parallelism": f"dp{a.gpus}_zero1_tp1_pp1", "micro_batch": a.micro_batch, "micro_batches": a.micro_batches},
        "gpu_launches": 0, "model_eval_wall_s": wall}))


# ----------------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: bench.py measures on a B200 (run it through gpurun)"}))
        return 1
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}"

    from simumax_b200.configs import get_simu_model_config
    from simumax_b200.core.config import ModelConfig
    from simumax_b200.ops import _lib
    from simumax_b200.runtime.trainer import LlamaDims, LlamaTrainer

    mc = ModelConfig.init_from_config_file(get_simu_model_config(a.model))
    dims = LlamaDims.from_model_config(mc)
    if a.layers:
        dims.layers = a.layers
    B, S, MB = a.micro_batch, a.seq_len, a.micro_batches
    tp = a.tp
    assert world % tp == 0, "--tp must divide --gpus"
    dp = world // tp
    if tp > 1:
        from simumax_b200.runtime.tp_trainer import TPLlamaTrainer
        tp_group = dp_group = None
        if dp > 1:                                   # rank layout as in the analytical model: tp fastest, then dp
            tp_groups = [dist.new_group(list(range(i * tp, (i + 1) * tp))) for i in range(dp)]
            dp_groups = [dist.new_group(list(range(j, world, tp))) for j in range(tp)]
            tp_group, dp_group = tp_groups[rank // tp], dp_groups[rank % tp]
        tr = TPLlamaTrainer(dims, B, S, tp_group=tp_group, dp_group=dp_group, seed=1234)
    else:
        tr = LlamaTrainer(dims, B, S, dp_group=dist.group.WORLD if world > 1 else None, attention_bwd=a.attention_bwd,
                          seed=1234)
    # distinct synthetic token batches per step and micro-batch, staged in pinned host memory (e2e) and on device
    n_sets = max(a.steps, 1)
    gen = torch.Generator().manual_seed(1000 + rank // tp)       # the ranks of one TP group share their replica's batch
    host = [[(torch.randint(0, mc.vocab_size, (B, S), generator=gen).pin_memory(),
              torch.randint(0, mc.vocab_size, (B, S), generator=gen).pin_memory()) for _ in range(MB)] for _ in range(n_sets)]
    dev = [[(t.cuda().view(-1), y.cuda().view(-1)) for t, y in step] for step in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2; the step itself streams >100 GB

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, label):
        for i in range(a.warmup):
            fn(i % n_sets)
        sampler = ClockSampler(local) if rank == 0 else None
        barrier()
        if sampler:
            sampler.start()
        l0 = _lib.launches()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        last = None
        for i in range(a.steps):
            flush.fill_(i & 0xFF)
            last = fn(i % n_sets)
        ev1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), wall_ms, clocks, _lib.launches() - l0, last

    dev_ms, _, clocks, launches, loss_t = timed(lambda i: tr.run_microbatches(dev[i]), "device")
    tokens_per_step = B * S * MB * dp
    value = tokens_per_step / (dev_ms / a.steps / 1e3)
    e2e = None
    if not a.skip_e2e:
        e_ms, e_wall, _, _, loss_f = timed(lambda i: tr.train_step(host[i]), "e2e")
        e2e = {"value": tokens_per_step / (max(e_ms, e_wall) / a.steps / 1e3), "unit": "tokens/s",
               "h2d_bytes_per_step": MB * 2 * B * S * 8, "d2h_bytes_per_step": 4, "ms_per_step": max(e_ms, e_wall) / a.steps,
               "api": f"{type(tr).__name__}.train_step(pinned host batches) -> float loss", "last_loss": loss_f}
    if rank == 0:
        mem = tr.memory_bytes()
        pred = predicted(a.model, world, S, B, MB, tp)
        flops_step = None
        out = {
            "metric": "llama3_8b_train_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": value / (BASELINE_TOK_S_PER_GPU * world), "dtype": "bf16",
            "data": "synthetic random tokens, random-init weights",
            "config": {"model": a.model, "global_batch": B * MB * dp, "seq_len": S,
                       "parallelism": f"dp{world}_zero1_tp1_pp1" if tp == 1 else f"dp{dp}_zero1_tp{tp}_sp_pp1 (fused AG->GEMM / GEMM->RS kernels)",
                       "micro_batch": B, "micro_batches": MB,
                       "layers": dims.layers, "padded_vocab": dims.vocab, "params": dims.total_numel(),
                       "optimizer": "AdamW fp32 master/m/v, fp32 main grads, grad clip 1.0", "recompute": "none",
                       "attention_bwd": getattr(tr, "attention_bwd", "smx"),
                       "l2": "256 MiB buffer rewritten before every step; a step touches >100 GB of weights/activations"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "predicted": pred, "memory_gb": {k: round(v / 2 ** 30, 3) for k, v in mem.items()},
            "last_loss": float(loss_t.item()),
        }
        if pred and "tokens_per_s" in pred:
            out["prediction_error_pct"] = 100.0 * (pred["tokens_per_s"] - value) / value
        try:     # achieved model FLOP rate against the driver-measured cuBLAS bf16 throughput (MEASURED_PEAKS.json)
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {"bf16_tflops": 1671.6, "bf16_tflops_sustained": 1391.5,
                                                                                    "note": "fallback: profiling recipe values"}
            mc.layer_num = dims.layers
            tf = value * mc.flops_per_token(S / 2) / 1e12 / world          # causal attention: half the S x S scores are computed
            out["roofline"] = {"model_tflops_per_gpu": tf, "model_tflops_per_gpu_full_attention_convention": value * mc.flops_per_token(S) / 1e12 / world,
                               "cublas_bf16_tflops_burst": peaks.get("bf16_tflops"),
                               "cublas_bf16_tflops_sustained": peaks.get("bf16_tflops_sustained"),
                               "fraction_of_sustained_cublas": tf / peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else None,
                               "fraction_of_nominal_2250": tf / 2250.0}
        except Exception as e:
            out["roofline"] = {"error": f"{type(e).__name__}: {e}"}
        if a.layers:
            out["invalid"] = "debug run with --layers override"
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main() or 0)
