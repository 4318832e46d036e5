"""Regression against numbers replayed from the reference code on its shipped B200 system config."""
import pytest

from helpers import MOE, build, headline, reference_system
from simumax_b200.core.utils import parse_bytes, parse_time_ms

GOLDEN = [
    # strategy, model, layers, model overrides, strategy overrides, iter ms, TGS/GPU, MFU, peak GiB per stage
    ("tp1_pp2_dp4_mbs1", "llama3-8b", None, {}, {}, 1006.6, 16275.99, 0.3723, [50.8854, 45.1637]),
    ("tp8_pp1_dp1_mbs1", "llama3-8b", None, {}, {}, 663.7697, 6170.82, 0.1412, [19.3472]),
    ("tp2_pp1_dp4_mbs1", "llama3-8b", None, {}, {}, 1050.0, 15603.38, 0.357, [43.6702]),
    ("tp2_pp1_dp4_mbs1", "llama3-70b", 12, {}, {}, 1342.3, 12206.18, 0.3946, [59.6541]),
    ("tp2_pp1_dp4_mbs1_selective_recompute", "llama3-70b", 12, {}, {}, 1667.0, 9828.45, 0.3178, [55.2449]),
    ("tp2_pp1_dp4_mbs1_full_recompute", "llama3-70b", 12, {}, {}, 1708.7, 9588.78, 0.31, [53.7449]),
    ("ep8_pp1_dp8_mbs1", "deepseekv2", 4, MOE, {}, 679.4734, 48225.58, 0.328, [45.8929]),
    ("ep4_pp2_dp4_mbs1", "deepseekv2", 4, MOE, {}, 485.0709, 33776.5, 0.2297, [30.3829, 44.9820]),
    ("tp2_pp1_dp4_mbs1", "llama3-70b", 12, {},
     dict(cross_entropy_loss_fusion=True, micro_batch_num=4, te_version="2.11.0", enable_recompute=False,
          enable_straggler_model=False), 698.8129, 11722.74, 0.379, [60.5134]),
    ("tp2_pp1_dp4_mbs1", "llama3-70b", 12, {},
     dict(cross_entropy_loss_fusion=True, micro_batch_num=4, te_version="2.11.0", enable_recompute=False,
          enable_straggler_model=False, cp_size=4, seq_len=32768), 1595.0, 10271.89, 0.4865, [68.1619]),
    ("tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt", "llama3-70b", 8, {}, {}, 725.5785, 11290.3, 0.2539,
     [43.2982, 28.9891, 26.8944, 39.1650]),
]


@pytest.mark.parametrize("case", GOLDEN, ids=[f"{c[1]}-{c[0]}-{i}" for i, c in enumerate(GOLDEN)])
def test_golden(case):
    strategy, model, layers, mover, sover, ms, tgs, mfu, peaks = case
    p = build(strategy, model, reference_system(), layers, mover, **sover)
    r = headline(p)
    assert not p.system.miss_efficiency
    assert parse_time_ms(r["duration_time_per_iter"]) == pytest.approx(ms, rel=2e-4)
    assert r["TGS_per_gpu"] == pytest.approx(tgs, rel=1e-4)
    assert r["mfu"] == pytest.approx(mfu, abs=6e-5)
    got = [r["peak_mem"]] if isinstance(r["peak_mem"], str) else list(r["peak_mem"].values())
    assert [parse_bytes(x) / 1024 ** 3 for x in got] == pytest.approx(peaks, abs=6e-5)


def test_golden_breakdown_llama3_8b_tp1_pp2():
    p = build("tp1_pp2_dp4_mbs1", "llama3-8b", reference_system())
    cost = p.analysis_cost().data
    first = p.model_chunk_dict["first_stage_chunk"].get_cost_info()
    assert first.fwd_compute_time == pytest.approx(30.0828, abs=1e-4)
    assert first.bwd_compute_time == pytest.approx(56.7868, abs=1e-4)
    assert p._compute_single_batch_fwd_bwd_time("last_stage_chunk", chunk=True) == pytest.approx(107.8141, abs=1e-3)
    dp = p._compute_dp_time("first_stage_chunk")["dense"]["details"]
    assert dp["reduce_scatter_time"] == pytest.approx(23.6484, abs=1e-4)
    assert dp["all_gather_time"] == pytest.approx(11.9864, abs=1e-4)
    assert p._compute_optim_time("first_stage_chunk")["optim_time"] == pytest.approx(21.6188, abs=1e-4)
    assert cost["flops_info"]["theory_flops"] == "843.3426 T"
    mem = p.analysis_mem().data
    assert mem["first_stage"]["model_mem"] == "33.6588 GB"
    assert mem["first_stage"]["fwd_activation_cache_per_micro_batch"] == "8.5039 GB"
    assert mem["last_stage"]["fwd_activation_cache_per_micro_batch"] == "10.5254 GB"


def test_a100_config():
    p = build("tp1_pp2_dp4_mbs1", "llama3-8b", reference_system("a100_pcie"))
    r = headline(p)
    assert parse_time_ms(r["duration_time_per_iter"]) == pytest.approx(5554.7, rel=1e-4)
    assert r["mfu"] == pytest.approx(0.4866, abs=1e-4)


def test_group_linear_mode_serial():
    """Legacy ``group_linear_mode='serial'`` (reference moe_module.py:878-900, 1101-1134): per-expert LinearCol children instead of one
    grouped GEMM.  Peak memory and the six table misses reproduce the reference's replay (32.0501 GB); the iteration is slower than
    the grouped mode.  (The reference's own chunk total leaves out the time of the row-linear children in this mode -- 3 x 20 x
    0.0194 ms forward per micro-batch on this config -- while pricing each of them; here they are counted.)"""
    from helpers import MOE, build, headline, reference_system
    sysc = reference_system()
    par = build("ep8_pp1_dp8_mbs1", "deepseekv2", sysc, layers=4, model_over=dict(MOE))
    ser = build("ep8_pp1_dp8_mbs1", "deepseekv2", sysc, layers=4, model_over=dict(MOE, group_linear_mode="serial"))
    rp, rs = headline(par), headline(ser)
    assert str(rs["peak_mem"]).startswith("32.0501"), rs["peak_mem"]
    kinds = [n.kind for n in list(ser.model_chunk_dict.values())[0].nodes]
    assert kinds.count("LinearCol") > kinds.count("LinearRow") + 100          # 3 MoE layers x 2 x 20 experts
    assert not any(k == "GroupLinear" for k in kinds)
    ms = lambda r: float(str(r["duration_time_per_iter"]).split()[0]) * (1e3 if str(r["duration_time_per_iter"]).strip().endswith(" s") else 1.0)
    assert ms(rs) > ms(rp)
