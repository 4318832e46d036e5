"""Large-scale strategy search CLI (MoE-friendly): sweeps tp x ep x pp with every recompute flavour for a world size
and global batch, keeps what fits under ``mem_gbs - gmi_error`` GiB, and dumps all candidates as CSV.

    python examples/search/llm_search.py --model deepseekv3 --system b200_nominal --world 2048 --gbs 16384 --ep 8 16 32 64
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from simumax_b200 import PerfLLM  # noqa: E402
from simumax_b200.configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config  # noqa: E402


def search(model, system, world, gbs, tp, ep, pp, kinds, margin_gib, dump):
    perf = PerfLLM()
    perf.configure(strategy_config=get_simu_strategy_config("tp1_pp2_dp4_mbs1"), model_config=get_simu_model_config(model),
                   system_config=get_simu_system_config(system))
    mc = perf.model_config
    mc.moe_pad_expert_input_to_capacity, mc.capacity = True, 1
    mc.padded_vocab_size, mc.make_vocab_size_divisible_by = True, 128
    perf.strategy.dispatch_probs = True
    results = {}
    best = perf.search_best_parallel_strategy(
        world_size=world, gmi_error=margin_gib, micro_batch_size=1, global_batch_size=gbs, all_search_result=results,
        tp_search_list=tp, ep_search_list=ep, pp_search_list=pp, recompute_search_type=kinds, use_reserved_memory=False,
        dump_path=dump or f"search_{mc.model_name}_{perf.system.sys_name}")
    return best, results


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="deepseekv2")
    ap.add_argument("--system", default="b200_nominal")
    ap.add_argument("--world", type=int, default=2048)
    ap.add_argument("--gbs", type=int, default=2048 * 8)
    ap.add_argument("--tp", type=int, nargs="*", default=[1])
    ap.add_argument("--ep", type=int, nargs="*", default=[8, 16, 32, 64])
    ap.add_argument("--pp", type=int, nargs="*", default=None)
    ap.add_argument("--recompute", nargs="*", default=["no_recompute", "full_block", "selective_recompute"])
    ap.add_argument("--margin-gib", type=int, default=6)
    ap.add_argument("--dump", default=None)
    a = ap.parse_args()
    best, _ = search(a.model, a.system, a.world, a.gbs, a.tp, a.ep, a.pp, a.recompute, a.margin_gib, a.dump)
    print(best)
