"""CPU-side calibration logic: shape enumeration, collective fits, system-config merge round trip."""
import json

import numpy as np
import pytest

from simumax_b200.calib.collectives import apply_to_system, fit_bw_latency, wire_bytes
from simumax_b200.calib.shapes import collect_requests, parse_key
from simumax_b200.calib.system_template import b200_template


def test_parse_key_and_requests():
    f = parse_key("b=1, m=4096, k=8192, n=10240, layout=TN, accumulate=False, out_dtype=bf16")
    assert f["m"] == "4096" and f["layout"] == "TN" and f["accumulate"] == "False"
    req = collect_requests([("llama3-8b", "tp1_pp1_dp8_mbs1", {}, {})])
    assert {"matmul", "sdp_fwd", "sdp_bwd"} <= set(req)
    keys = list(req["matmul"])
    assert any("layout=NT" in k and "accumulate=True" in k and "out_dtype=fp32" in k for k in keys)
    r = next(iter(req["matmul"].values()))
    assert r.flops == 2.0 * r.geti("b") * r.geti("m") * r.geti("k") * r.geti("n")


@pytest.mark.parametrize("op,n", [("all_reduce", 8), ("all_gather", 4), ("reduce_scatter", 2), ("all2all", 8), ("p2p", 2)])
def test_fit_recovers_synthetic_curve(op, n):
    gbps, eff, lat = 900.0, 0.62, 3.5
    sizes = [1 << k for k in range(16, 29)]
    scale = {"all_reduce": 2}.get(op, 1)
    hops = (n - 1) * scale if op != "p2p" else 1
    t = [wire_bytes(op, s, n) / (gbps * 1024 ** 3 * eff) * 1e6 + lat * hops for s in sizes]
    f = fit_bw_latency(op, n, sizes, t, gbps)
    assert f["efficient_factor"] == pytest.approx(eff, rel=1e-6)
    assert f["latency_us"] == pytest.approx(lat, rel=1e-6)
    assert f["max_rel_err"] < 1e-9


def test_fitted_system_reproduces_measurement(tmp_path):
    """Write a fit into a system JSON, load it through SystemConfig and check compute_net_op_time returns the data."""
    from simumax_b200.core.config import SystemConfig
    sizes = [1 << k for k in range(18, 28)]
    fits = {}
    truth = {}
    for n in (2, 4, 8):
        fits[n] = {}
        for op in ("all_reduce", "all_gather", "reduce_scatter", "all2all"):
            eff, lat = 0.5 + 0.03 * n, 2.0 + 0.1 * n
            scale = 2 if op == "all_reduce" else 1
            t = [wire_bytes(op, s, n) / (900 * 1024 ** 3 * eff) * 1e6 + lat * (n - 1) * scale for s in sizes]
            fits[n][op] = fit_bw_latency(op, n, sizes, t, 900.0)
            truth[(n, op)] = (sizes[5], t[5])
    system = apply_to_system(b200_template("unit"), fits)
    p = tmp_path / "sys.json"
    p.write_text(json.dumps(system))
    s = SystemConfig.init_from_config_file(str(p))
    for (n, op), (size, t_us) in truth.items():
        got_ms = s.compute_net_op_time(op, size, n, "high_intra_node", comm_stage="tp")
        assert got_ms * 1e3 == pytest.approx(t_us, rel=2e-3), (n, op)


def test_merge_writes_loadable_config(tmp_path):
    from simumax_b200.calib.calibrate import merge
    from simumax_b200.core.config import SystemConfig
    key = "b=1, m=4096, k=4096, n=6144, layout=TN, accumulate=False, out_dtype=bf16"
    ops = {"ops": {"matmul": {key: {"ms": 0.15, "tflops": 1374.0, "flops": 2.06e11, "efficiency": 0.61}}},
           "bandwidth": {"default": {"gbs": 5200.0}, "ce_fusion": {"gbs": 4100.0}}, "meta": {"device": "unit"}}
    (tmp_path / "ops.json").write_text(json.dumps(ops))
    out = tmp_path / "cal.json"
    merge(str(tmp_path / "ops.json"), [], str(out), sys_name="unit_cal")
    s = SystemConfig.init_from_config_file(str(out))
    assert s.accelerator.op["matmul"].accurate_efficient_factor[key] == pytest.approx(0.61)
    t = s.compute_op_accuracy_time("matmul", 2.06e11, key)
    assert t == pytest.approx(2.06e11 / (2250e12 * 0.61) * 1e3)
    assert s.accelerator.bandwidth["ce_fusion"].efficient_factor == pytest.approx(4100e9 / (8000 * 1024 ** 3))


def test_fused_exposure_tables_feed_the_tp_overlap_model():
    """calib/fused.py turns fused_check results into {tp_size: exposure} tables the strategy accepts."""
    import glob
    import os
    from helpers import build
    from simumax_b200.calib.fused import exposure_from_fused_check, strategy_overrides
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "profiles", "fused_ops_2gpu_v4.json")) + glob.glob(os.path.join(root, "profiles", "fused_ops_4gpu_v3.json"))
                   + glob.glob(os.path.join(root, "profiles", "fused_ops_8gpu_v1.json")))
    table = exposure_from_fused_check(files)
    assert set(table) == {2, 4, 8} and all(0.0 < v["rs"] < 1.0 and v["rs_cases"] == 3 for v in table.values())
    over = strategy_overrides(table)
    ms = lambda v: float(str(v).split()[0])
    base = build("tp2_pp1_dp4_mbs1", "llama3-8b", "b200_nominal", 4)
    tab = build("tp2_pp1_dp4_mbs1", "llama3-8b", "b200_nominal", 4, **over)
    flat = build("tp2_pp1_dp4_mbs1", "llama3-8b", "b200_nominal", 4, tp_overlap_model=True,
                 tp_overlap_exposed_ag=over["tp_overlap_exposed_ag"]["2"], tp_overlap_exposed_rs=over["tp_overlap_exposed_rs"]["2"])
    t = [ms(p.analysis_cost().data["breakdown_result"]["intra_exposed_time"]) for p in (base, tab, flat)]
    assert t[1] < t[0] and abs(t[1] - t[2]) < 1e-9
