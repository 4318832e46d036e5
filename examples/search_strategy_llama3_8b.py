"""Strategy search walk-through on Llama-3-8B / 8 GPUs:
  1. best (micro_batch_size, micro_batch_num) split of a fixed global batch that fits in memory,
  2. best (tp, pp) under a memory margin, without recompute,
  3. the plain grid searcher (tuning.StrategySearcher) for comparison.
"""
import os
import sys
from pprint import pprint

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _cases import build, default_system  # noqa: E402

from simumax_b200.configs import get_simu_model_config, get_simu_system_config  # noqa: E402
from simumax_b200.tuning import StrategySearcher  # noqa: E402


def fixed_global_batch(perf, gbs=32):
    mbs, mbc, peaks, costs = perf.search_max_micro_batch_size_fixed_gbs(
        pp_size=perf.strategy.pp_size, dp_size=perf.strategy.dp_size, global_batch_size=gbs, gmi_error=10,
        use_reserved_memory=True, save_all=False, verbose=False)
    pprint({"micro_batch_size": mbs[0], "micro_batch_num": mbc[0], "peak_mem_gib_by_stage": peaks[0],
            "iter_time": costs[0].data["duration_time_per_iter"], "tokens/s/gpu": costs[0].data["throughput_per_accelerator"]})


def parallel_strategy(perf, gbs=32):
    everything = {}
    best = perf.search_best_parallel_strategy(
        world_size=8, gmi_error=10, micro_batch_size=1, global_batch_size=gbs, all_search_result=everything,
        tp_search_list=[1, 2, 4], ep_search_list=[1], pp_search_list=[1, 2], recompute_search_type=["no_recompute"],
        use_reserved_memory=True, dump_path=None, verbose=False)
    pprint(best)
    print(f"({len(everything)} candidates evaluated)")


def main():
    perf = build("llama3_8b_tp1_pp2")
    perf.strategy.enable_recompute = False
    print("== fixed global batch ==")
    fixed_global_batch(perf)
    print("== parallel strategy ==")
    parallel_strategy(perf)
    print("== grid searcher ==")
    s = StrategySearcher(get_simu_model_config("llama3-8b"), get_simu_system_config(default_system()))
    for hit in s.search(topk=3):
        pprint(hit.summary())


if __name__ == "__main__":
    main()
