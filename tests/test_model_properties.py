"""Structural invariants of the op IR that must hold for any config (no reference data needed)."""
import os

import pytest

from helpers import MOE, build

CASES = [("tp2_pp1_dp4_mbs1", "llama3-8b", 4, {}, {}),
         ("tp2_pp1_dp4_mbs1_selective_recompute", "llama3-70b", 3, {}, {"recompute_layer_num": 3}),
         ("tp2_pp1_dp4_mbs1_full_recompute", "llama3-70b", 3, {}, {"recompute_layer_num": 2}),
         ("ep8_pp1_dp8_mbs1", "deepseekv2", 3, MOE, {}),
         ("ep8_pp1_dp8_mbs1", "deepseekv3", 3, MOE, {"dispatch_probs": True}),
         ("tp2_pp1_dp4_mbs1", "llama3-70b", 2, {}, {"cp_size": 2, "seq_len": 8192}),
         ("tp1_pp2_dp4_mbs1", "mixtral-8x7b", 4, {}, {"ep_size": 4, "fp8": True})]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[1]}-{i}" for i, c in enumerate(CASES)])
def test_chunk_invariants(case):
    strategy, model, layers, mover, sover = case
    p = build(strategy, model, "b200_nominal", layers, mover, **sover)
    for name, chunk in p.model_chunk_dict.items():
        pk = p.pp_state_peak_point[name]
        assert pk.peak_mem >= pk.activation_mem_cache >= 0
        total = sum(n.cost.fwd_compute_time for n in chunk.nodes)
        assert chunk.get_cost_info().fwd_compute_time == pytest.approx(total)
        assert all(n.cost.fwd_compute_time >= 0 and n.cost.bwd_compute_time >= 0 for n in chunk.nodes)
        # recompute never reduces time, and cached bytes with recompute never exceed the no-recompute cache
        gemm = chunk.get_all_gemm_cost_info()
        assert len(gemm["Module"]) == 3 * sum(1 for n in chunk.nodes if n.gemm is not None)      # router rows included, like the reference
        info = chunk.analysis_op_info()
        assert len(info["op"]) >= 2 * len(chunk.nodes)
    mem = p.analysis_mem().data
    cost = p.analysis_cost().data
    assert cost["raw"]["duration_time_per_iter_ms"] > 0 and 0 < cost["mfu"] < 1


def test_recompute_trades_memory_for_time():
    base = build("tp2_pp1_dp4_mbs1", "llama3-70b", "b200_nominal", 4)
    full = build("tp2_pp1_dp4_mbs1_full_recompute", "llama3-70b", "b200_nominal", 4, recompute_layer_num=4)
    t0 = base.analysis_cost().data["raw"]["duration_time_per_iter_ms"]
    t1 = full.analysis_cost().data["raw"]["duration_time_per_iter_ms"]
    m0 = base.analysis_mem().data["raw"]["peak_mem"]
    m1 = full.analysis_mem().data["raw"]["peak_mem"]
    assert t1 > t0 and m1 < m0


def test_tp_shards_weights():
    a = build("tp2_pp1_dp4_mbs1", "llama3-8b", "b200_nominal", 2)
    b = build("tp4_pp1_dp2_mbs1", "llama3-8b", "b200_nominal", 2)
    wa = a.model_chunk_dict["first_stage_chunk"].get_model_info().dense_weight_bytes
    wb = b.model_chunk_dict["first_stage_chunk"].get_model_info().dense_weight_bytes
    assert wa > wb > wa / 2.2


def test_result_files(tmp_path):
    import json
    p = build("tp1_pp2_dp4_mbs1", "llama3-8b", "b200_nominal", 4)
    p.analysis(str(tmp_path), console_log=False)
    for f in ("compute_result.json", "mem_result.json", "net_info.json", "base_info.json", "strategy_config.json",
              "system_config.json", "model_config.json", "model_arch"):
        assert (tmp_path / f).exists(), f
    assert "duration_time_per_iter" in json.load(open(tmp_path / "compute_result.json"))
    p.export_pp_schedule_trace(str(tmp_path / "pp.json"))
    tr = json.load(open(tmp_path / "pp.json"))
    assert tr["displayTimeUnit"] == "ms" and any(e.get("ph") == "X" for e in tr["traceEvents"])
    g = p.capture(str(tmp_path))
    assert len(g.nodes) > 10 and (tmp_path / "model_graph.json").exists()
    # the reference's ONNX-style schema (core/graph.py:96-124); both pipeline chunks in one graph, call_idx restarts per chunk
    gj = json.load(open(tmp_path / "model_graph.json"))
    assert list(gj)[:6] == ["name", "nodes", "inputs", "outputs", "initializers", "value_info"] and gj["name"] == "model_graph"
    first = gj["nodes"][0]
    assert first["name"] == "embedding_0" and first["inputs"] == ["input_0"] and first["outputs"] == ["output_1"]
    assert set(first) == {"name", "op_type", "enable_recompute", "is_variance_node", "inputs", "outputs", "call_idx", "attributes", "domain"}
    assert gj["value_info"]["output_1"] == {"shape": [1, 4096, 4096], "dtype": "bf16"}
    assert sum(n["call_idx"] == 0 for n in gj["nodes"]) == 2 and sum(t.startswith("input_") for n in gj["nodes"] for t in n["inputs"]) == 2
    from simumax_b200.core.graph import visualize_with_graphviz
    assert os.path.exists(visualize_with_graphviz(str(tmp_path / "model_graph.json"), str(tmp_path / "g")))


def test_model_arch_is_the_nested_module_view(tmp_path):
    """``model_arch``: torch-style nested blocks, identical layers folded into ``(a->b)``, and under every module its time /
    FLOP / static-memory totals (reference MetaModule.__repr__, base_struct.py:1004-1068)."""
    import json
    import re
    p = build("tp2_pp1_dp4_mbs1", "llama3-8b", "b200_nominal", 4)
    p.analysis(save_path=str(tmp_path), console_log=False)
    text = (tmp_path / "model_arch").read_text()
    assert text.startswith("{'first_stage_chunk': LLMModel(\n  (0): Embedding(hidden_size=4096,vocab_size=")
    assert "  (1->4): LLMBlock(\n" in text                       # four identical layers, one entry
    assert "      (0): LinearCol(input_size=4096,output_size=3072,use_bias=False,enable_recompute=False, TP=2)\n" in text
    assert "(1): CoreAttention(head_size=128,head_num=16.0,kv_head_num=4.0,use_math_sdp=False,use_flash_sdp=True," in text
    assert text.count("\t1. cost: (total_time=") == text.count("\t2. memory: (d_w=") == 16   # 12 leaves + 3 composites + root
    assert json.loads((tmp_path / "base_info.json").read_text())["arch"] == text
    # a composite's totals are the sums of its children's
    chunk = p.model_chunk_dict["first_stage_chunk"]
    root_total = float(re.findall(r"total_time=([0-9.]+) ms", text)[-1])
    assert root_total == pytest.approx(chunk.get_cost_info().all_time, abs=0.005)
    assert "self.layer_0.mlp.linear_fc1" in chunk.summary()      # the flat per-op table is still available
