#!/usr/bin/env python
"""What the analytical model predicts for one routed-expert MoE block (router, permutation + dispatch all-to-all, grouped FC1,
SwiGLU, grouped FC2, combine all-to-all + unpermutation; forward + backward) -- the quantity ``tools/moe_check.py`` measures with
``runtime/moe_layer.py``.  CPU only.

    python tools/b200/predict_moe_block.py [--model deepseekv2] [--ep 8] [--system b200_sm100a_calibrated] [--measured-ms 5.99]
"""
import argparse
import contextlib
import io
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def predict(model="deepseekv2", ep=8, system="b200_sm100a_calibrated", seq=4096, mbs=1):
    from simumax_b200 import PerfLLM
    from simumax_b200.configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config
    from simumax_b200.core.config import StrategyConfig
    warnings.filterwarnings("ignore")
    st = StrategyConfig.init_from_config_file(get_simu_strategy_config("ep8_pp1_dp8_mbs1"))
    st.world_size = st.ep_size = ep
    st.seq_len, st.micro_batch_size, st.dispatch_probs = seq, mbs, True
    p = PerfLLM()
    with contextlib.redirect_stdout(io.StringIO()):
        p.configure(strategy_config=st, model_config=get_simu_model_config(model), system_config=get_simu_system_config(system))
        p.model_config.layer_num = max(2, p.model_config.dense_layers + 1)
        p.model_config.moe_pad_expert_input_to_capacity, p.model_config.capacity = True, 1     # balanced routing, capacity factor 1
        p.run_estimate()
    layer = p.model_config.layer_num - 1
    rows = []
    for n in p.model_chunk_dict["first_stage_chunk"].nodes:
        if f".layer_{layer}.mlp." in n.name and "shared" not in n.name.lower() and n.kind != "AddModule":
            c = n.cost
            rows.append({"op": n.name.split(".mlp.")[1], "kind": n.kind, "fwd_ms": c.fwd_compute_time,
                         "bwd_ms": c.bwd_grad_act_time + c.bwd_grad_w_time, "fwd_net_ms": c.fwd_net_time,
                         "bwd_net_ms": c.bwd_grad_act_net_time + c.bwd_grad_w_net_time})
    total = sum(r["fwd_ms"] + r["bwd_ms"] + r["fwd_net_ms"] + r["bwd_net_ms"] for r in rows)
    return {"model": model, "ep": ep, "system": system, "tokens_per_rank": seq * mbs, "rows": rows, "fwd_bwd_ms": total,
            "efficiency_misses": {k: len(v) for k, v in p.system.miss_efficiency.items()}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="deepseekv2")
    ap.add_argument("--ep", type=int, default=8)
    ap.add_argument("--system", default="b200_sm100a_calibrated")
    ap.add_argument("--measured-ms", type=float, default=None)
    a = ap.parse_args()
    r = predict(a.model, a.ep, a.system)
    for row in r["rows"]:
        print(f"{row['op']:28s} {row['kind']:16s} fwd {row['fwd_ms']:.4f}  bwd {row['bwd_ms']:.4f}  net {row['fwd_net_ms'] + row['bwd_net_ms']:.4f} ms")
    msg = f"predicted forward + backward: {r['fwd_bwd_ms']:.3f} ms"
    if a.measured_ms:
        msg += f"   measured {a.measured_ms:.3f} ms   error {100 * (r['fwd_bwd_ms'] - a.measured_ms) / a.measured_ms:+.1f} %"
    print(msg, "  misses:", json.dumps(r["efficiency_misses"]))
