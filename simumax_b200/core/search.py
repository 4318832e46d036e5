"""Search API: largest micro-batch, fixed-GBS micro-batch sweep, recompute-layer binary search, selective
recompute combos and the tp x ep x pp parallel-strategy grid.

Signatures and result dictionaries follow the reference (perf_llm.py:3080-3578) so that search scripts and
CSV consumers keep working; candidates are evaluated on the flat op IR, where identical transformer layers
are built and costed once per chunk and cloned (``models.llm.StageChunk`` block cache), which is this
framework's replacement for the reference's per-unit search profile cache.
"""
from __future__ import annotations

import math
import os
from copy import deepcopy
from typing import List



def merge_dict(cur: dict, merged: dict) -> dict:
    if len(merged) == 0:
        for k, v in cur.items():
            merged[k] = [v]
    else:
        for k, v in cur.items():
            merged.setdefault(k, []).append(v)
    return merged


def _cached_search(fn):
    """Run a search entry point with the block profile cache switched on (outermost call owns it; SIMU_SEARCH_CACHE=0
    disables it for A/B runs).  Reference: the search-only CachedChunkProfile machinery, perf_llm.py:837-1379."""
    import functools
    import os

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        if getattr(self, "_block_cache", None) is not None or os.environ.get("SIMU_SEARCH_CACHE", "1") == "0":
            return fn(self, *args, **kwargs)
        from ..models.llm import BlockProfileCache
        self._block_cache = BlockProfileCache()
        try:
            return fn(self, *args, **kwargs)
        finally:
            self.last_search_cache_stats = {"hits": self._block_cache.hits, "misses": self._block_cache.misses}
            self._block_cache = None
    return wrapper


class SearchMixin:
    # ------------------------------------------------------------------ helpers
    def _peak_key(self, use_reserved_memory: bool) -> str:
        return "peak_mem_with_reserved" if use_reserved_memory else "peak_mem"

    def _evaluate(self, peak_key: str):
        self.run_estimate()
        mem = self.analysis_mem()
        cost = self.analysis_cost()
        peaks = self.get_pp_stage_peak_mem(mem, peak_key, toG=True)
        return mem, cost, peaks

    def log_available_strategy(self, mfu, peak_mem):
        if getattr(self, "_search_verbose", True):
            print(f"Find result  parallelism={self.strategy.parallelism}, pp_num_layers={self.get_pp_num_layers()}, "
                  f"recompute={self.strategy.recompute_status},mfu={mfu} gbs={self.strategy.global_batch_size} "
                  f"peak_cached_mem_bytes={peak_mem}GB", flush=True)

    def get_pp_num_layers(self):
        per = math.ceil(self.model_config.layer_num / self.strategy.pp_size)
        if self.strategy.pp_size > 1:
            return f"[{per}]x{self.strategy.pp_size - 1} + [{self.strategy.num_layers_in_last_pipeline_stage}]"
        return [self.model_config.layer_num]

    def dump_paralism_and_recompute_perf(self, mem_result, cost_result):
        md, cd = mem_result.data, cost_result.data
        dtype = "fp8" if self.strategy.fp8 else "bf16"
        pick = lambda key: md[key] if key in md else {s: v[key] for s, v in md.items()}
        return {"model_name": self.model_config.model_name, "param_num": cd["param_numel_info"]["all"],
                "system": self.system.sys_name,
                "parallelism": f"{dtype}.dense{self.model_config.dense_layers}.{self.strategy.parallelism}",
                "recompute_status": self.strategy.recompute_status, "mfu": cd["mfu_6nd_with_attn"],
                "TFLOPS": cd["throughput per GPU (TFLOP/s/GPU)"], "TGS_per_gpu": cd["throughput_per_accelerator"],
                "iter_time": cd["duration_time_per_iter"], "straggle_ratio": cd["straggle_ratio"],
                "activations_ratio": cd["param_numel_info"]["activations_ratio"], "peak_mem": pick("peak_mem"),
                "peak_mem_with_reserved": pick("peak_mem_with_reserved")}

    def dump_paralism_and_recompute_bw_perf(self, mem_result, cost_result):
        perf = self.dump_paralism_and_recompute_perf(mem_result, cost_result)
        perf["comm_bw_info"] = str(deepcopy(self.system.real_comm_bw))
        return perf

    # ------------------------------------------------------------------ micro-batch searches
    @_cached_search
    def search_max_micro_batch_size(self, micro_batch_num=None):
        """Binary search of the largest micro-batch size that fits ``accelerator.mem_gbs`` (fixed micro-batch count)."""
        left, right = 1, 2 ** 16
        budget = self.system.accelerator.mem_gbs * 1024 ** 3
        keep = (self.strategy.micro_batch_size, self.strategy.micro_batch_num)
        self.strategy.micro_batch_num = self.strategy.pp_size * 16 if micro_batch_num is None else micro_batch_num
        peak = 0
        mbs = left
        try:
            while left < right:
                mbs = left + ((right - left) >> 1)
                self.strategy.micro_batch_size = mbs
                self.run_estimate()
                peak = max(self.get_pp_stage_peak_mem(self.analysis_mem(), "peak_mem", False).values())
                if peak > budget:
                    right = mbs
                else:
                    left = mbs + 1
        finally:
            self.strategy.micro_batch_size, self.strategy.micro_batch_num = keep
        if getattr(self, "_search_verbose", True):
            print(f"cur micro_batch_size: {mbs}, micro_batch_num: {micro_batch_num}")
            print(f"Peak cached memory: {peak / 1024 ** 3: .2f} GB")
        return left - 1, peak

    @_cached_search
    def search_max_micro_batch_size_fixed_gbs(self, pp_size, dp_size, global_batch_size, memory_utils=1.0, gmi_error=6,
                                              use_reserved_memory=True, save_all=True, verbose: bool = True):
        """All (micro_batch_size, micro_batch_num) splits of a fixed global batch that fit in memory, largest first."""
        margin = gmi_error * 1024 ** 3
        key = self._peak_key(use_reserved_memory)
        budget = self.system.accelerator.mem_gbs * 1024 ** 3 * memory_utils
        keep = (self.strategy.micro_batch_size, self.strategy.micro_batch_num, getattr(self, "_search_verbose", True))
        self._search_verbose = verbose
        sizes, nums, peaks, costs = [], [], [], []
        try:
            for mbs in range(global_batch_size - 1, 0, -1):
                if global_batch_size % (mbs * dp_size) != 0:
                    continue
                mbn = global_batch_size // (mbs * dp_size)
                if global_batch_size % mbs != 0 or mbn < pp_size:
                    continue
                self.strategy.micro_batch_num, self.strategy.micro_batch_size = mbn, mbs
                self.run_estimate()
                mem = self.analysis_mem()
                peak = max(self.get_pp_stage_peak_mem(mem, key, False).values())
                if peak + margin <= budget:
                    cost = self.analysis_cost()
                    peak_g = self.get_pp_stage_peak_mem(mem, key, toG=True)
                    if not save_all:
                        return [mbs], [mbn], [peak_g], [cost]
                    sizes.append(mbs); nums.append(mbn); peaks.append(peak_g); costs.append(cost)
            return sizes, nums, peaks, costs
        finally:
            self.strategy.micro_batch_size, self.strategy.micro_batch_num, self._search_verbose = keep

    # ------------------------------------------------------------------ recompute searches
    @_cached_search
    def search_best_selective_recompute(self, use_reserved_memory, gmi_error, best_mfu=None, all_search_result=None,
                                        save_path=None):
        if self.strategy.megatron_recompute:
            raise NotImplementedError("search does not support megatron_recompute yet; "
                                      "please evaluate megatron_recompute strategies explicitly")
        self.strategy.recompute_granularity = "selective_recompute"
        budget_g = self.system.accelerator.mem_gbs - gmi_error
        key = self._peak_key(use_reserved_memory)
        best = {}
        best_mfu = -1 if best_mfu is None else best_mfu
        combos = [dict(mla_rms_recompute=True, attn_recompute=True, mlp_rms_recompute=True, mlp_recompute=True),
                  dict(mla_rms_recompute=True, attn_recompute=True, mlp_rms_recompute=False, mlp_recompute=False),
                  dict(mla_rms_recompute=False, attn_recompute=False, mlp_rms_recompute=True, mlp_recompute=True)]
        for combo in combos:
            for k, v in combo.items():
                setattr(self.strategy, k, v)
            mem, cost, peaks = self._evaluate(key)
            peak_g = max(peaks.values())
            if peak_g <= budget_g:
                cur = self.dump_paralism_and_recompute_bw_perf(mem, cost)
                if cur["mfu"] > best_mfu:
                    best_mfu, best = cur["mfu"], cur
                    self.log_available_strategy(cost.data["mfu"], peak_g)
                    if save_path is not None:
                        self._dump_memory_and_cost(mem, cost, save_path)
                if all_search_result is not None:
                    merge_dict(cur, all_search_result)
        return best

    @_cached_search
    def search_best_recompute_layer_num(self, layer_num, use_reserved_memory: bool, gmi_error: int, best_mfu,
                                        all_search_result: dict, save_path=None):
        """Binary search of the smallest number of recomputed layers per stage that fits (fewest = fastest)."""
        budget_g = self.system.accelerator.mem_gbs - gmi_error
        key = self._peak_key(use_reserved_memory)
        best = {}
        left, right = 0, math.ceil(layer_num / self.strategy.pp_size)
        keep = self.strategy.recompute_layer_num
        try:
            while left <= right:
                n = (left + right) // 2
                assert n <= math.ceil(layer_num / self.strategy.pp_size)
                self.strategy.recompute_layer_num = n
                mem, cost, peaks = self._evaluate(key)
                peak_g = max(peaks.values())
                if peak_g > budget_g:
                    left = n + 1
                else:
                    right = n - 1
                    if cost.data["mfu"] >= best_mfu:
                        best_mfu = cost.data["mfu"]
                        best = self.dump_paralism_and_recompute_bw_perf(mem, cost)
                        self.log_available_strategy(cost.data["mfu"], peak_g)
                        if save_path is not None:
                            self._dump_memory_and_cost(mem, cost, save_path)
                    if all_search_result is not None:
                        merge_dict(self.dump_paralism_and_recompute_bw_perf(mem, cost), all_search_result)
        finally:
            self.strategy.recompute_layer_num = keep
        return best

    @_cached_search
    def search_best_strategy_no_recompute(self, gmi_error, use_reserved_memory, best_mfu, all_search_result,
                                          save_path=None):
        self.strategy.recompute_granularity = None
        self.strategy.recompute_layer_num = 0
        budget_g = self.system.accelerator.mem_gbs - gmi_error
        mem, cost, peaks = self._evaluate(self._peak_key(use_reserved_memory))
        peak_g = max(peaks.values())
        best = {}
        if peak_g <= budget_g:
            cur = self.dump_paralism_and_recompute_bw_perf(mem, cost)
            if all_search_result is not None:
                merge_dict(cur, all_search_result)
            if cost.data["mfu"] > best_mfu:
                best = cur
                self.log_available_strategy(cost.data["mfu"], peak_g)
                if save_path is not None:
                    self._dump_memory_and_cost(mem, cost, save_path)
        return best

    # ------------------------------------------------------------------ parallel strategy grid
    @_cached_search
    def search_best_parallel_strategy(self, world_size: int, gmi_error: int, micro_batch_size: int,
                                      global_batch_size: int, all_search_result: dict, tp_search_list: List = None,
                                      ep_search_list: List = None, pp_search_list: List = None, use_etp: bool = False,
                                      recompute_search_type=("no_recompute", "full_block", "selective_recompute"),
                                      use_reserved_memory: bool = True, dump_path: str = None, verbose: bool = True):
        """Grid over tp x ep x pp (uneven last stage allowed) x recompute flavour at fixed micro/global batch;
        keeps the feasible candidate with the highest MFU and optionally dumps CSV tables."""
        if self.strategy.megatron_recompute:
            raise NotImplementedError("search does not support megatron_recompute yet; "
                                      "please evaluate megatron_recompute strategies explicitly")
        if not isinstance(recompute_search_type, (list, tuple)):
            recompute_search_type = [recompute_search_type]
        m = self.model_config
        layer_num = m.layer_num
        tp_search_list = tp_search_list or ([1, 2, 4, 8] if m.model_type == "dense" else [1])
        ep_search_list = ep_search_list or ([1, 2, 4, 8] if m.model_type == "moe" else [1])
        pp_search_list = pp_search_list or list(range(1, layer_num + 1))
        if m.expert_num == 1:
            ep_search_list = [1]
        else:
            tp_search_list = [1]
        keep_verbose, keep_strategy = getattr(self, "_search_verbose", True), self.strategy
        self._search_verbose = verbose
        cost_path = f"{dump_path}/best_strategy_costs" if dump_path is not None else None
        best, best_mfu = {}, -1

        def candidate(tp, ep, etp, pp, last_layers):
            cand = deepcopy(keep_strategy)
            cand.world_size, cand.tp_size, cand.ep_size, cand.pp_size, cand.etp_size = world_size, tp, ep, pp, etp
            cand.num_layers_in_first_pipeline_stage = None
            cand.num_layers_in_last_pipeline_stage = last_layers
            origin = self.strategy
            try:
                cand.sanity_check()
                self.strategy = cand
                self._cross_sanity_check()
                return cand
            except (AssertionError, ValueError, ZeroDivisionError) as exc:
                if verbose:
                    print("skip invalid strategy by sanity_check:", {"world_size": world_size, "tp_size": tp,
                          "cp_size": cand.cp_size, "ep_size": ep, "etp_size": etp, "pp_size": pp, "reason": str(exc)})
                return None
            finally:
                self.strategy = origin

        def last_stage_layers(pp):
            if pp <= 1:
                return True, None
            per = math.ceil(layer_num / pp)
            last = layer_num - per * (pp - 1)
            return per > 0 and last > 0, last

        if verbose:
            print(f"Start search strategy for world_size={world_size}, model_type={m.model_type}, "
                  f"model_name={m.model_name}, system={self.system.sys_name}")
            print(f"- tp_search_list={tp_search_list}, ep_search_list={ep_search_list}, pp_search_list={pp_search_list}")
            print(f"- layer_num={layer_num}")
        try:
            for tp in tp_search_list:
                for ep in ep_search_list:
                    for pp in pp_search_list:
                        ok, last = last_stage_layers(pp)
                        cand = candidate(tp, ep, tp if use_etp else 1, pp, last) if ok else None
                        if cand is None:
                            continue
                        self.strategy = cand
                        if global_batch_size % (cand.dp_size * micro_batch_size) != 0:
                            continue
                        mbn = global_batch_size // (cand.dp_size * micro_batch_size)
                        cand.micro_batch_num, cand.micro_batch_size = mbn, micro_batch_size
                        if micro_batch_size == 0 or mbn == 0:
                            continue
                        for kind in recompute_search_type:
                            keep_var = cand.recompute_variance
                            try:
                                if kind == "no_recompute":
                                    cand.recompute_variance = True
                                    found = self.search_best_strategy_no_recompute(
                                        gmi_error=gmi_error, use_reserved_memory=use_reserved_memory,
                                        best_mfu=best_mfu, all_search_result=all_search_result)
                                elif kind == "full_block":
                                    cand.recompute_granularity = "full_block"
                                    cand.recompute_variance = False
                                    found = self.search_best_recompute_layer_num(
                                        layer_num=layer_num, use_reserved_memory=use_reserved_memory,
                                        gmi_error=gmi_error, best_mfu=best_mfu, all_search_result=all_search_result,
                                        save_path=cost_path)
                                elif kind == "layer_only":
                                    found = self.search_best_recompute_layer_num(
                                        layer_num=layer_num, use_reserved_memory=use_reserved_memory,
                                        gmi_error=gmi_error, best_mfu=best_mfu, all_search_result=all_search_result,
                                        save_path=cost_path)
                                elif kind == "selective_recompute":
                                    cand.recompute_granularity = "selective_recompute"
                                    cand.recompute_layer_num = math.ceil(layer_num / pp)
                                    found = self.search_best_selective_recompute(
                                        use_reserved_memory=use_reserved_memory, gmi_error=gmi_error, best_mfu=best_mfu,
                                        all_search_result=all_search_result, save_path=cost_path)
                                else:
                                    raise NotImplementedError(f"recompute strategy {kind} not implemented")
                            finally:
                                cand.recompute_variance = keep_var
                            if found and "mfu" in found:
                                best, best_mfu = found, found["mfu"]
                            self.system.reset_record_info()
            if dump_path is not None and best:
                import pandas as pd
                os.makedirs(dump_path, exist_ok=True)
                row = dict(best)
                for k in ("peak_mem", "peak_mem_with_reserved"):
                    if isinstance(row.get(k), dict):
                        row[k] = str(row[k])
                stem = (f"{dump_path}/{m.model_name}_{self.system.sys_name}_seqlen{self.strategy.seq_len}_"
                        f"worldsize{self.strategy.world_size}_gbs{self.strategy.global_batch_size}")
                pd.DataFrame(row, index=[0]).to_csv(f"{stem}_best_strategy.csv")
                if all_search_result:
                    n = max(len(v) for v in all_search_result.values())
                    table = {k: (v + [None] * (n - len(v))) for k, v in all_search_result.items()}
                    df = pd.DataFrame({k: [str(x) if isinstance(x, dict) else x for x in v] for k, v in table.items()})
                    df.sort_values(by="mfu", ascending=False).to_csv(f"{stem}_all_search_strategies.csv")
            return best
        finally:
            self._search_verbose = keep_verbose
            self.strategy = keep_strategy
