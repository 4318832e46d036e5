"""Exhaustive grid search over strategy knobs (the simple sibling of ``PerfLLM.search_best_parallel_strategy``).

For every point of the user grid x every legal (pp, ep, etp) factorisation of the world it finds the largest micro-batch
that fits, evaluates the iteration, and keeps the top-k by MFU.  Unlike the reference's legacy tool
(/root/reference/simumax/tuning/strategy_searcher.py:143-213, which still treats the tuple returned by
``search_max_micro_batch_size`` as a scalar) this one is exercised by the test-suite.
"""
from __future__ import annotations

import contextlib
import io
import itertools
from copy import deepcopy
from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional

from ..core.config import ModelConfig, StrategyConfig, SystemConfig
from ..core.perf_llm import PerfLLM

CANDIDATE_STRATEGY: Dict[str, list] = {
    "seq_len": [4096],
    "world_size": [8],
    "micro_batch_size": [1],
    "micro_batch_num": [8],
    "dtype": ["bf16"],
    "tp_size": [1, 2, 4, 8],
    "enable_sequence_parallel": [True],
    "zero_state": [1],
    "enable_recompute": [False, True],
    "recompute_granularity": ["full_block"],
}


@dataclass
class SearchHit:
    strategy: StrategyConfig
    mfu: float
    tokens_per_gpu_s: float
    iter_time: object
    peak_mem_bytes: float

    def summary(self) -> dict:
        s = self.strategy
        return {"tp": s.tp_size, "pp": s.pp_size, "ep": s.ep_size, "etp": s.etp_size, "mbs": s.micro_batch_size,
                "mbc": s.micro_batch_num, "recompute_layers": getattr(s, "recompute_layer_num", 0) if s.enable_recompute else 0,
                "mfu": self.mfu, "tgs": self.tokens_per_gpu_s, "iter_time": self.iter_time,
                "peak_gb": self.peak_mem_bytes / 1024 ** 3}


class StrategySearcher:
    def __init__(self, model_config, system_config, verbose: bool = False):
        self.model_config = model_config if isinstance(model_config, ModelConfig) else ModelConfig.init_from_config_file(model_config)
        self.system_config = system_config if isinstance(system_config, SystemConfig) else SystemConfig.init_from_config_file(system_config)
        self.verbose = verbose
        self.hits: List[SearchHit] = []
        self.skipped: List[tuple] = []

    # ---- enumeration
    def parallel_factorisations(self, world_size: int, tp_size: int) -> Iterator[dict]:
        """Legal (pp, ep, etp) for this world/tp: layers divisible by pp, experts divisible by ep, world divisible by all."""
        m = self.model_config
        rest = world_size // tp_size
        for pp in range(1, rest + 1):
            if rest % pp or m.layer_num % pp:
                continue
            if getattr(m, "expert_num", 1) in (0, 1, None):
                yield {"pp_size": pp, "ep_size": 1, "etp_size": 1}
                continue
            etp = 1
            while etp <= self.system_config.num_per_node:
                for ep in range(1, m.expert_num + 1):
                    if m.expert_num % ep == 0 and (world_size // pp) % (ep * etp) == 0:
                        yield {"pp_size": pp, "ep_size": ep, "etp_size": etp}
                etp *= 2

    def generate_grid(self, candidates: Dict[str, list], recompute_buckets: int = 4) -> List[dict]:
        out = []
        keys = list(candidates)
        for values in itertools.product(*(candidates[k] for k in keys)):
            base = dict(zip(keys, values))
            if base["world_size"] % base["tp_size"]:
                continue
            for par in self.parallel_factorisations(base["world_size"], base["tp_size"]):
                p = {**base, **par}
                if not p.get("enable_recompute"):
                    out.append(p)
                    continue
                layers = self.model_config.layer_num // p["pp_size"]
                step = max(1, layers // recompute_buckets)
                for n in range(layers, 0, -step):
                    out.append({**p, "recompute_layer_num": n})
        return out

    # ---- evaluation
    def evaluate(self, params: dict, max_micro_batch: bool = True) -> Optional[SearchHit]:
        strategy = StrategyConfig(**params)
        perf = PerfLLM()
        perf._search_verbose = False
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink if not self.verbose else contextlib.nullcontext()) if not self.verbose else contextlib.nullcontext():
            perf.configure(strategy_config=strategy, model_config=deepcopy(self.model_config),
                           system_config=deepcopy(self.system_config))
            if max_micro_batch:
                mbs, _peak = perf.search_max_micro_batch_size(micro_batch_num=strategy.micro_batch_num)
                if mbs < 1:
                    return None
                perf.strategy.micro_batch_size = mbs
            perf.run_estimate()
            cost = perf.analysis_cost().data
            mem = perf.analysis_mem()
            peak = max(perf.get_pp_stage_peak_mem(mem, "peak_mem", False).values())
        if peak > self.system_config.accelerator.mem_gbs * 1024 ** 3:
            return None
        return SearchHit(deepcopy(perf.strategy), cost["mfu"], cost["throughput_per_accelerator"],
                         cost["duration_time_per_iter"], peak)

    def search(self, topk: int = 5, candidate_strategy: Optional[Dict[str, list]] = None, max_micro_batch: bool = True
               ) -> List[SearchHit]:
        self.hits, self.skipped = [], []
        for params in self.generate_grid(candidate_strategy or CANDIDATE_STRATEGY):
            try:
                hit = self.evaluate(params, max_micro_batch)
            except (AssertionError, ValueError, NotImplementedError) as err:
                self.skipped.append((params, f"{type(err).__name__}: {err}"))
                continue
            if hit is None:
                self.skipped.append((params, "does not fit in memory"))
            else:
                self.hits.append(hit)
        self.hits.sort(key=lambda h: h.mfu, reverse=True)
        return self.hits[:topk]
