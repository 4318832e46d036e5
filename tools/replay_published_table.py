"""Replay the reference's published B200 table (docs/b200/b200_release_v1.2_summary.json, 43 rows: Llama-3-70B / 405B, DeepSeek-V2 / V3,
TP / PP / EP / CP 32 K - 128 K) on the reference's shipped system config with BOTH implementations and compare with the recorded
`perf_ms` / `perf_alloc_gib` columns.  CPU only.
    python tools/replay_published_table.py [out.json]        (REF = reference checkout, default /root/reference)
Recipe per row (reference tools/b200/run_megatron_perf_real_pipeline.py: dense_case / moe_case / configure_perf, `_cef` = fused CE):
sizes from the case name, te_version 2.11.0, no recompute, no straggler model, fused cross-entropy, vocab padded to 128; MoE rows pad the
expert input to capacity 1 and dispatch probabilities; DeepSeek-V3 rows use dense_layers = 1."""
import contextlib
import io
import json
import os
import re
import sys
import warnings

warnings.filterwarnings("ignore")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
CFG = f"{REF}/configs"


def recipe(row):
    name = row["case"]
    g = lambda key, default=1: int(m.group(1)) if (m := re.search(rf"_{key}(\d+)", name)) else default
    tp, pp, ep, cp, seq = g("tp"), g("pp"), g("ep"), g("cp"), g("seq", 4096)
    moe = name.startswith("deepseek")
    if moe:
        base = "ep8_pp1_dp8_mbs1" if pp == 1 else "ep4_pp2_dp4_mbs1"
    elif cp > 1:
        base = "tp2_pp1_dp4_mbs1" if tp == 2 else "tp1_pp1_dp8_mbs1"
    else:
        base = {(1, 2): "tp1_pp2_dp4_mbs1", (2, 1): "tp2_pp1_dp4_mbs1", (4, 1): "tp4_pp1_dp2_mbs1", (8, 1): "tp8_pp1_dp1_mbs1"}[(tp, pp)]
    st = dict(world_size=8, tp_size=tp, pp_size=pp, ep_size=ep, cp_size=cp, micro_batch_size=1, micro_batch_num=row["mbc"], seq_len=seq,
              te_version="2.11.0", enable_recompute=False, enable_straggler_model=False, cross_entropy_loss_fusion=True)
    mo = dict(layer_num=row["layer_num"], padded_vocab_size=True, make_vocab_size_divisible_by=128)
    if moe:
        st["dispatch_probs"] = True
        mo.update(moe_pad_expert_input_to_capacity=True, capacity=1)
        if name.startswith("deepseekv3"):
            mo["dense_layers"] = 1
    model = row["model"]
    if not os.path.exists(f"{CFG}/models/{model}.json"):                 # the table says "llama3-405b", the shipped file carries a suffix
        cands = sorted(f[:-5] for f in os.listdir(f"{CFG}/models") if f.startswith(model))
        model = cands[0] if cands else model
    return model, base, st, mo


def run(which, row):
    if which == "reference":
        from simumax.core.perf_llm import PerfLLM
    else:
        from simumax_b200 import PerfLLM
    model, base, st, mo = recipe(row)
    p = PerfLLM()
    with contextlib.redirect_stdout(io.StringIO()):
        p.configure(strategy_config=f"{CFG}/strategy/{base}.json", model_config=f"{CFG}/models/{model}.json",
                    system_config=f"{CFG}/system/b200_bf16_ceperm.json")
        for k, v in st.items():
            setattr(p.strategy, k, v)
        for k, v in mo.items():
            setattr(p.model_config, k, v)
        p.run_estimate()
        cost = p.analysis_cost().data
        mem = p.analysis_mem()
        stages = p.get_pp_stage_peak_mem(mem, "peak_mem", toG=True)
    t = cost["duration_time_per_iter"]
    ms = float(t.split()[0]) * (1000.0 if t.strip().endswith(" s") else 1.0)
    return round(ms, 2), round(max(float(v) for v in stages.values()), 2), {k: len(v) for k, v in p.system.miss_efficiency.items()}


def vpp_preview(which="simumax_b200"):
    """The sync-VPP preview row of the reference's release notes (docs/release_v1.2.md:33-35): llama3_70b_l8_tp1_pp4_vp2_dp2_mbc8_sync,
    perf 661.21 ms, simulator trace end 663.29 ms.  Returns (perf ms, simulator end ms)."""
    import tempfile
    if which == "reference":
        from simumax.core.perf_llm import PerfLLM
    else:
        from simumax_b200 import PerfLLM
    p = PerfLLM()
    with contextlib.redirect_stdout(io.StringIO()):
        p.configure(strategy_config=f"{CFG}/strategy/tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt.json", model_config=f"{CFG}/models/llama3-70b.json",
                    system_config=f"{CFG}/system/b200_bf16_ceperm.json")
        for k, v in dict(world_size=8, tp_size=1, pp_size=4, micro_batch_num=8, interleaving_size=2, te_version="2.11.0", pp_comm_async=False,
                         enable_straggler_model=False, cross_entropy_loss_fusion=True).items():
            setattr(p.strategy, k, v)
        for k, v in dict(layer_num=8, padded_vocab_size=True, make_vocab_size_divisible_by=128).items():
            setattr(p.model_config, k, v)
        p.run_estimate()
        t = p.analysis_cost().data["duration_time_per_iter"]
        with tempfile.TemporaryDirectory() as d:
            sim = p.simulate(d)
    ms = float(t.split()[0]) * (1000.0 if t.strip().endswith(" s") else 1.0)
    end = sim["end_time_ms"] if isinstance(sim, dict) and "end_time_ms" in sim else None
    return round(ms, 2), (round(end, 2) if end is not None else None)


def main():
    rows = json.load(open(f"{REF}/docs/b200/b200_release_v1.2_summary.json"))["rows"]
    out, same_impl, same_rec_ms, same_rec_mem = [], 0, 0, 0
    for row in rows:
        try:
            new = run("simumax_b200", row)
            ref = run("reference", row)
        except Exception as e:                                  # a row this script cannot express: report, keep going
            out.append({"case": row["case"], "error": f"{type(e).__name__}: {e}"[:200]})
            print(f"{row['case']:58s} ERROR {out[-1]['error']}")
            continue
        rec = (row["perf_ms"], row["perf_alloc_gib"])
        ok_impl = new[:2] == ref[:2]
        ok_ms = abs(new[0] - rec[0]) <= max(0.011, 2e-4 * rec[0])
        ok_mem = abs(new[1] - rec[1]) <= 0.011
        same_impl += ok_impl
        same_rec_ms += ok_ms
        same_rec_mem += ok_mem
        out.append({"case": row["case"], "this_repo_ms": new[0], "reference_code_ms": ref[0], "recorded_ms": rec[0], "this_repo_gib": new[1],
                    "reference_code_gib": ref[1], "recorded_gib": rec[1], "efficiency_misses": new[2], "real_ms": row.get("real_ms")})
        print(f"{row['case']:58s} {new[0]:10.2f} {ref[0]:10.2f} {rec[0]:10.2f}   {new[1]:7.2f} {ref[1]:7.2f} {rec[1]:7.2f}  "
              f"{'=' if ok_impl else 'IMPL-DIFF'} {'' if ok_ms and ok_mem else 'vs-record: ' + ('ms ' if not ok_ms else '') + ('mem' if not ok_mem else '')}")
    n = len([o for o in out if "error" not in o])
    summary = {"rows": len(rows), "replayed": n, "identical_to_reference_code": same_impl, "perf_ms_equals_recorded": same_rec_ms,
               "perf_alloc_equals_recorded": same_rec_mem}
    summary["sync_vpp_preview_row"] = {"this_repo_perf_ms_and_simulator_end_ms": vpp_preview(), "recorded": [661.21, 663.29]}
    print(json.dumps(summary))
    if len(sys.argv) > 1:
        json.dump({"summary": summary, "rows": out}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
