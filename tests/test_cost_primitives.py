import math

import pytest

from simumax_b200.core.config import SystemConfig
from simumax_b200.core.schedule import one_f_one_b, one_f_one_b_simple, straggler_increase_ratio

from helpers import reference_system
from simumax_b200.configs import get_simu_system_config


def nominal():
    return SystemConfig.init_from_config_file(get_simu_system_config("b200_nominal"))


def test_net_time_formula_matches_closed_form():
    s = nominal()
    size, n = 160e6, 4
    t = s.compute_net_op_time("reduce_scatter", size, n, "high_intra_node", comm_stage="dp_cp")
    op = s.networks["high_intra_node"].op["reduce_scatter"]
    wire = size * 1 + (size / n) * (-1)
    lat = op.latency_us * (n - 1) * 1 + op.fixed_latency_us
    assert t == pytest.approx(wire / (900 * 1024 ** 3 * op.efficient_factor) * 1e3 + lat / 1e3, rel=1e-12)
    assert s.compute_net_op_time("all_reduce", size, 1, "high_intra_node") == 0
    rec = s.real_comm_bw["reduce_scatter"]["dp_cp"]
    assert rec["comm_num"] == 4 and rec["real_bw"] == pytest.approx(900 * op.efficient_factor)


def test_reference_b200_collective_record():
    s = SystemConfig.init_from_config_file(reference_system())
    s.compute_net_op_time("reduce_scatter", 160e6, 4, "high_intra_node", comm_stage="dp_cp")
    rec = s.real_comm_bw["reduce_scatter"]["dp_cp"]
    assert rec["latency"] == pytest.approx(24.0975, abs=1e-4)
    assert rec["real_bw"] == pytest.approx(605.79, abs=1e-2)


def test_roofline_and_shape_table():
    s = nominal()
    s.accelerator.op["matmul"].accurate_efficient_factor = {"k1": 0.5}
    t_hit = s.compute_op_accuracy_time("matmul", 2.25e15, "k1")
    t_miss = s.compute_op_accuracy_time("matmul", 2.25e15, "other")
    assert t_hit == pytest.approx(2000.0) and t_miss == pytest.approx(1000.0 / 0.55)
    assert "matmul" in s.hit_efficiency and "shape=other" in s.miss_efficiency["matmul"]
    assert s.compute_end2end_time(3.0, 5.0) == 5.0
    s.accelerator.mode = "only_compute"
    assert s.compute_end2end_time(3.0, 5.0) == 3.0 and s.compute_end2end_time(0, 5.0) == 5.0


def test_1f1b_properties():
    ph = lambda f, b, p: {"fwd_recv": p, "fwd_compute": f, "fwd_send": p, "bwd_recv": p, "bwd_compute": b, "bwd_send": p}
    # single stage: pure sum
    one = {"fwd_recv": 0, "fwd_compute": 2.0, "fwd_send": 0, "bwd_recv": 0, "bwd_compute": 3.0, "bwd_send": 0}
    assert one_f_one_b(1, 5, [one]) == pytest.approx(25.0)
    # uniform stages without comm: classic bubble (pp - 1) * (f + b)
    for pp, mbc in ((2, 4), (4, 8), (4, 3)):
        # (a p2p time of exactly 0 removes the rendezvous ops altogether, so use a vanishing one)
        stages = [ph(1.0, 2.0, 1e-9) for _ in range(pp)]
        assert one_f_one_b(pp, mbc, stages) == pytest.approx((mbc + pp - 1) * 3.0, rel=1e-6)
        assert one_f_one_b_simple(pp, mbc, [1.0] * pp, [2.0] * pp) == pytest.approx((mbc + pp - 1) * 3.0)
    # p2p only adds time and total is monotone in micro-batch count
    stages = [ph(1.0, 2.0, 0.1) for _ in range(4)]
    t8, t9 = one_f_one_b(4, 8, stages), one_f_one_b(4, 9, stages)
    assert t9 > t8 > (8 + 3) * 3.0


def test_straggler_ratio():
    assert straggler_increase_ratio(1) == 1.0
    L = math.log2(8)
    assert straggler_increase_ratio(8) == pytest.approx(1 + L / (L + 1) * 0.09 * math.sqrt(L))


def test_runtime_semantics_knobs_are_off_by_default_and_monotone():
    """tp_overlap_model / fused_optimizer_model describe this repo's fused runtime; default off = reference semantics."""
    from helpers import build

    def run(**over):
        p = build("tp2_pp1_dp4_mbs1", "llama3-8b", "b200_nominal", 8, **over)
        c = p.analysis_cost().data
        ms = lambda v: float(str(v).split()[0])
        return (ms(c["duration_time_per_iter"]), ms(c["breakdown_result"]["intra_exposed_time"]),
                ms(c["breakdown_result"]["optim_time"]), p.analysis_mem().data)

    base = run()
    assert run(tp_overlap_model=False, fused_optimizer_model=False)[:3] == base[:3]
    tp = run(tp_overlap_model=True)
    assert tp[1] < 0.7 * base[1] and tp[2] == base[2] and tp[0] < base[0]
    opt = run(fused_optimizer_model=True)
    assert opt[2] < 0.6 * base[2] and opt[1] == base[1]
    # the fused runtime saves the gathered activation instead of re-gathering it for the wgrad: more cached bytes
    from simumax_b200.core.utils import parse_bytes
    peak = lambda d: max(parse_bytes(v["peak_mem"]) for v in d.values()) if "peak_mem" not in d else parse_bytes(d["peak_mem"])
    assert peak(tp[3]) > peak(base[3])


def test_context_parallel_flavours_extension():
    """cp_comm_type: "a2a" is the reference's formal path; "all_gather" (finished here) and "p2p" (ring) are extensions: same
    attention work per rank, different collectives."""
    from helpers import build

    def run(**over):
        p = build("tp2_pp1_dp4_mbs1", "llama3-70b", "b200_nominal", 8, seq_len=32768, cp_size=4, world_size=32, **over)
        c = p.analysis_cost().data
        ms = lambda v: float(str(v).split()[0]) * (1e3 if str(v).split()[1] == "s" else 1.0)
        return ms(c["duration_time_per_iter"]), ms(c["breakdown_result"]["intra_exposed_time"]), c["flops_info"]["model_flops"]

    a2a, ag, ring = run(cp_comm_type="a2a"), run(cp_comm_type="all_gather"), run(cp_comm_type="p2p")
    hidden = run(cp_comm_type="p2p", cp_p2p_exposed=0.0)
    assert a2a[2] == ag[2] == ring[2]                     # identical model FLOPs
    assert ring[1] < a2a[1] and hidden[1] < ring[1]       # neighbour transfers are cheaper than the all-to-alls; hiding them helps
    assert all(0.8 * a2a[0] < x[0] < 1.2 * a2a[0] for x in (ag, ring, hidden))


def test_kernel_launch_overhead_extension(tmp_path):
    """accelerator.kernel_launch_overhead_us (default 0 = reference): a constant per kernel-launching op phase; survives the
    config round trip and scales with the op count, not with the problem size."""
    import json
    from helpers import build
    from simumax_b200.configs import get_simu_system_config
    from simumax_b200.core.config import SystemConfig
    base = json.load(open(get_simu_system_config("b200_nominal")))
    assert "kernel_launch_overhead_us" not in base["accelerator"]
    base["accelerator"]["kernel_launch_overhead_us"] = 5.0
    base["sys_name"] = "b200_launch5"
    path = tmp_path / "sys.json"
    path.write_text(json.dumps(base))
    assert SystemConfig.init_from_config_file(str(path)).to_config_dict()["accelerator"]["kernel_launch_overhead_us"] == 5.0
    ms = lambda p: float(str(p.analysis_cost().data["duration_time_per_iter"]).split()[0])
    plain = build("tp8_pp1_dp1_mbs1", "llama3-8b", "b200_nominal", 8, micro_batch_num=4)
    slow = build("tp8_pp1_dp1_mbs1", "llama3-8b", str(path), 8, micro_batch_num=4)
    ops = sum(1 for n in plain.model_chunk_dict["first_stage_chunk"].nodes for ph in ("fwd", "bwd_grad_act", "bwd_grad_w")
              if getattr(n.compute, {"fwd": "fwd_flops", "bwd_grad_act": "bwd_grad_act_flops", "bwd_grad_w": "bwd_grad_w_flops"}[ph])
              or getattr(n.compute, {"fwd": "fwd_accessed_mem", "bwd_grad_act": "bwd_grad_act_accessed_mem", "bwd_grad_w": "bwd_grad_w_accessed_mem"}[ph]))
    extra = ms(slow) - ms(plain)
    assert abs(extra - 4 * ops * 5e-3) < 0.02 * extra and extra > 1.0
