from helpers import build


def test_fixed_gbs_and_strategy_search(tmp_path):
    p = build("tp1_pp2_dp4_mbs1", "llama3-8b", "b200_nominal", 8, run=False)
    sizes, nums, peaks, costs = p.search_max_micro_batch_size_fixed_gbs(pp_size=2, dp_size=4, global_batch_size=32,
                                                                        gmi_error=10, save_all=True, verbose=False)
    assert sizes and all(s * n * 4 == 32 for s, n in zip(sizes, nums))
    assert sizes == sorted(sizes, reverse=True)
    assert (p.strategy.micro_batch_size, p.strategy.micro_batch_num) == (1, 8)      # restored
    allr = {}
    best = p.search_best_parallel_strategy(world_size=8, gmi_error=10, micro_batch_size=1, global_batch_size=32,
                                           all_search_result=allr, tp_search_list=[1, 2], ep_search_list=[1],
                                           pp_search_list=[1, 2, 4], dump_path=str(tmp_path), verbose=False)
    assert best and best["mfu"] == max(allr["mfu"])
    assert any(f.name.endswith("_best_strategy.csv") for f in tmp_path.iterdir())
    assert p.strategy.pp_size == 2                                                    # original strategy restored


def test_recompute_layer_search_picks_minimal_feasible():
    p = build("tp2_pp1_dp4_mbs1_full_recompute", "llama3-70b", "b200_nominal", 8, run=False)
    p.system.accelerator.mem_gbs = 60
    res = {}
    best = p.search_best_recompute_layer_num(layer_num=8, use_reserved_memory=False, gmi_error=0, best_mfu=-1,
                                             all_search_result=res)
    assert best, "some recompute depth must fit in 60 GB"
    assert best["mfu"] > 0 and len(res["mfu"]) >= 1


def test_max_micro_batch_size():
    p = build("tp2_pp1_dp4_mbs1", "llama3-8b", "b200_nominal", 8, run=False)
    mbs, peak = p.search_max_micro_batch_size(micro_batch_num=2)
    assert mbs >= 1 and peak > 0


def test_search_cache_is_transparent(monkeypatch):
    """Block profile cache (reference: CachedChunkProfile, perf_llm.py:41-252) must not change a single reported number,
    including the achieved-bandwidth / efficiency bookkeeping, and must actually be hit."""
    import json

    def run(model, strategy, layers, **kw):
        p = build(strategy, model, "b200_nominal", layers, run=False)
        allr = {}
        best = p.search_best_parallel_strategy(world_size=8, gmi_error=10, micro_batch_size=1, global_batch_size=32,
                                               all_search_result=allr, verbose=False, **kw)
        return json.dumps([best, allr], sort_keys=True, default=str), getattr(p, "last_search_cache_stats", None)

    cases = [("llama3-70b", "tp2_pp1_dp4_mbs1", 16, dict(tp_search_list=[2, 4, 8], ep_search_list=[1], pp_search_list=[1, 2],
                                                       recompute_search_type=["no_recompute", "full_block", "selective_recompute"])),
             ("deepseekv2", "ep8_pp1_dp8_mbs1", 6, dict(tp_search_list=[1], ep_search_list=[4, 8], pp_search_list=[1, 2],
                                                     recompute_search_type=["no_recompute", "full_block"]))]
    for model, strategy, layers, kw in cases:
        monkeypatch.setenv("SIMU_SEARCH_CACHE", "0")
        plain, stats0 = run(model, strategy, layers, **kw)
        monkeypatch.setenv("SIMU_SEARCH_CACHE", "1")
        cached, stats1 = run(model, strategy, layers, **kw)
        assert stats0 is None and stats1["hits"] > stats1["misses"] > 0
        assert plain == cached
