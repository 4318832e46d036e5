import json

import pytest

from simumax_b200.configs import (get_simu_model_config, get_simu_strategy_config, get_simu_system_config, list_models,
                                  list_strategies, list_systems)
from simumax_b200.core.config import ModelConfig, StrategyConfig, SystemConfig


@pytest.mark.parametrize("name", list_models())
def test_model_roundtrip(name):
    m = ModelConfig.init_from_config_file(get_simu_model_config(name))
    d = m.to_dict()
    assert d["hidden_size"] == m.hidden_size and d["param_numel"] == m.param_numel > 0
    assert m.model_type in ("dense", "moe")
    assert m.flops_per_token(4096) > 6 * m.activated_param_numel * 0.5
    json.dumps(d)


@pytest.mark.parametrize("name", list_strategies())
def test_strategy_roundtrip_and_sanity(name):
    import warnings
    warnings.filterwarnings("ignore")
    s = StrategyConfig.init_from_config_file(get_simu_strategy_config(name))
    s.sanity_check()
    assert s.dp_size * s.tp_size * s.cp_size * s.pp_size == s.world_size
    assert s.global_batch_size == s.micro_batch_size * s.micro_batch_num * s.dp_size
    s2 = StrategyConfig.init_from_format_strings(s.parallelism)
    assert (s2.tp_size, s2.pp_size, s2.ep_size, s2.seq_len, s2.micro_batch_num) == \
           (s.tp_size, s.pp_size, s.ep_size, s.seq_len, s.micro_batch_num)


def test_unknown_key_rejected():
    with pytest.raises(TypeError):
        StrategyConfig.init_from_dict({"seq_len": 1, "not_a_field": 2})


def test_vocab_padding():
    m = ModelConfig.init_from_config_file(get_simu_model_config("llama3-8b"))
    m.maybe_pad_vocab_size(8)
    assert m.vocab_size % (128 * 8) == 0 and m.vocab_size >= m.orig_vocab_size == 128257
    m.maybe_pad_vocab_size(1)
    assert m.vocab_size == 128384


def test_te_version_gates_and_recompute_parsing():
    s = StrategyConfig(seq_len=4096, micro_batch_size=1, micro_batch_num=1, te_version="2.11.0")
    assert s.te_dummy_wgrad_memory_enabled and s.te_cp_a2a_saves_pre_posta2a_output
    s.te_version = "2.2"
    assert not s.te_dummy_wgrad_memory_enabled
    s = StrategyConfig(seq_len=4096, micro_batch_size=1, micro_batch_num=1, recompute_granularity="selective_recompute",
                       recompute_layer_num=2, attn_recompute=True, mlp_recompute=True)
    a, m = s.parse_attention_recompute(0), s.parse_mlp_recompute(0)
    assert a.core_attn_recompute and not a.out_recompute and m.linear_recompute and not m.pre_mlp_norm_recompute
    assert not s.parse_attention_recompute(2).core_attn_recompute
    s.megatron_recompute = True
    with pytest.raises(AssertionError):
        s.sanity_check()


@pytest.mark.parametrize("name", list_systems())
def test_system_loads_and_serialises(name):
    s = SystemConfig.init_from_config_file(get_simu_system_config(name))
    again = SystemConfig.init_from_dict(s.to_config_dict())
    assert again.to_config_dict() == s.to_config_dict()
    assert s.compute_mem_access_time("default", 0) == 0
    assert s.compute_mem_access_time("default", 1 << 30) > 0


def test_command_line_front_end(tmp_path, capsys):
    """python -m simumax_b200 perf / simulate / trace / list (the reference ships example scripts only)."""
    import json
    from simumax_b200.__main__ import main
    assert main(["perf", "-m", "llama3-8b", "-s", "tp2_pp1_dp4_mbs1", "--system", "b200_nominal", "--layers", "4",
                 "--set", "micro_batch_num=4", "--set", "model.vocab_size=32000"]) == 0
    head = json.loads(capsys.readouterr().out)
    assert head["mfu"] > 0 and "mbc4" in head["parallelism"] and head["tokens_per_s_per_gpu"] > 0
    out = tmp_path / "sim"
    assert main(["simulate", "-m", "llama3-8b", "-s", "tp1_pp2_dp4_mbs1", "--system", "b200_nominal", "--layers", "4", "--out", str(out)]) == 0
    sim = json.loads(capsys.readouterr().out)
    assert sim["lanes"] == 2 and sim["simulated_end_ms"] > 0 and (out / "tracing_logs.json").exists()
    assert main(["trace", str(out / "log.log"), str(out / "again.json")]) == 0 and (out / "again.json").exists()
    capsys.readouterr()
    assert main(["list"]) == 0
    assert "llama3-8b" in json.loads(capsys.readouterr().out)["models"]


def test_debug_points_write_cost_log(tmp_path, monkeypatch):
    """configure(debug_points=[...]) -> <TMP_PATH>/cost_log.json with the F / B / W / collective times of the matching ops."""
    import json
    import simumax_b200.core.perf_llm as pl
    from helpers import build
    monkeypatch.setattr(pl, "TMP_PATH", str(tmp_path))
    p = build("tp2_pp1_dp4_mbs1", "llama3-8b", "b200_nominal", 2, run=False)
    p.debug_points = ["self.layer_1.mlp.linear_fc1"]
    p.debug_points_last_stage = ["parallel_ce"]
    p.run_estimate()
    log = json.load(open(tmp_path / "cost_log.json"))
    (key, rec), = [(k, v) for k, v in log.items() if "linear_fc1" in k]
    assert "first_stage_chunk" in key and rec["cost_F"] > 0 and rec["cost_W"] > 0 and rec["net_F"] > 0
