import pytest

from simumax_b200.core.tensor import BPE, FakeTensor, Float8Tensor, TensorSize


def test_tensor_algebra():
    t = TensorSize((4096, 2, 4096))
    assert t.numel() == 4096 * 2 * 4096 and t.get_memory_size() == t.numel() * BPE["bf16"]
    assert t.view(4096, 2, 32, -1).shape == [4096, 2, 32, 128]
    assert t.transpose(0, 1).shape == [2, 4096, 4096] and t.permute(2, 0, 1).shape == [4096, 4096, 2]
    assert t.unsqueeze(-1).shape == [4096, 2, 4096, 1] and t.unsqueeze(0).squeeze(0) == t
    assert TensorSize((1, 8)).expand(4, -1).shape == [4, 8]
    parts = t.split([1024, 3072], dim=-1)
    assert [p.shape[-1] for p in parts] == [1024, 3072] and TensorSize.cat(parts, -1) == t
    assert t.new_with_dim(1, 8).size(1) == 8 and t.size() == [4096, 2, 4096]
    assert FakeTensor is TensorSize and t.view(-1).parents == (t,)
    with pytest.raises(ValueError):
        t.view(7, -1)
    with pytest.raises(IndexError):
        t.size(5)
    f8 = Float8Tensor((16, 16))
    assert f8.dtype == "fp8" and f8.get_memory_size() == 256


def test_grid_searcher_finds_fitting_strategies():
    from simumax_b200.configs import get_simu_model_config, get_simu_system_config
    from simumax_b200.tuning import StrategySearcher
    s = StrategySearcher(get_simu_model_config("llama3-8b"), get_simu_system_config("b200_nominal"))
    grid = {"seq_len": [4096], "world_size": [8], "micro_batch_size": [1], "micro_batch_num": [8], "dtype": ["bf16"],
            "tp_size": [1, 2], "enable_sequence_parallel": [True], "zero_state": [1], "enable_recompute": [False]}
    cands = s.generate_grid(grid)
    assert {(c["tp_size"], c["pp_size"]) for c in cands} >= {(1, 1), (1, 2), (1, 4), (1, 8), (2, 1), (2, 2), (2, 4)}
    top = s.search(topk=3, candidate_strategy=grid)
    assert len(top) == 3 and top[0].mfu >= top[1].mfu >= top[2].mfu
    budget = s.system_config.accelerator.mem_gbs * 1024 ** 3
    assert all(h.peak_mem_bytes <= budget and h.strategy.micro_batch_size >= 1 for h in top)


def test_dualpipe_what_if_is_wired_to_the_model():
    """PerfLLM.estimate_dualpipe(): the DualPipe / ZB1P closed forms (pp_simu/utils.py) fed with the evaluated stage chunk."""
    from helpers import MOE, build
    p = build("ep4_pp2_dp4_mbs1", "deepseekv2", "b200_nominal", 8, model_over=MOE, pp_size=4, world_size=32, micro_batch_num=16)
    d = p.estimate_dualpipe()
    assert d["pp_size_used"] == 4 and d["f_ms"] > 0 and d["b_ms"] > d["f_ms"] > d["w_ms"] > 0
    assert d["f_and_b_overlapped_ms"] <= d["f_and_b_serial_ms"]
    assert d["bubble_dualpipe"] < d["bubble_zb1p"] < d["bubble_1f1b"]
    assert d["mfu_dualpipe"] > d["mfu_zb1p"] > d["mfu_1f1b"] > 0
    # default = the slowest stage: its 1F1B closed form agrees with the schedule evaluator up to p2p and stage imbalance
    c = p.analysis_cost().data
    sched = c["raw"]["schedule_time_ms"]
    assert d["chunk"] == "last_stage_chunk" and abs(d["duration_1f1b"] - sched) / sched < 0.05
    assert p.estimate_dualpipe("first_stage_chunk")["duration_1f1b"] < d["duration_1f1b"]
