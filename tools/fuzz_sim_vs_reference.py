"""Differential fuzzing of the discrete-event simulator against the reference's: random (model, strategy) pairs in merged-lane
mode (one simulated thread per pipeline stage), end time and per-rank static / peak memory must be identical or both must reject.

World mode (one thread per rank) is excluded on purpose: the reference's world-wide optimizer all-reduce expects ``pp_size``
arrivals there (base_struct.py:2192-2195), so it does not actually synchronise the ranks; this implementation does
(tools/compare_sim.py case 12 shows the resulting 1e-4 ms difference).

    python tools/fuzz_sim_vs_reference.py [n_cases] [seed]
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from compare_sim import MOE, run  # noqa: E402

DENSE = [("llama3-8b", "b200_bf16_ceperm"), ("llama3-70b", "b200_bf16_ceperm"), ("llama2-tiny", "a100_pcie"), ("llama2-7b", "b200_bf16_ceperm")]
MOES = [("deepseekv2", "b200_bf16_ceperm"), ("mixtral-8x7b", "b200_bf16_ceperm"), ("deepseekv2-lite", "b200_bf16_ceperm")]


def sample(rng):
    moe = rng.random() < 0.35
    model, system = rng.choice(MOES if moe else DENSE)
    base = rng.choice(["ep8_pp1_dp8_mbs1", "ep4_pp2_dp4_mbs1"] if moe else
                      ["tp1_pp2_dp4_mbs1", "tp2_pp1_dp4_mbs1", "tp2_pp1_dp4_mbs1_full_recompute", "tp2_pp1_dp4_mbs1_selective_recompute",
                       "tp4_pp1_dp2_mbs1", "tp1_pp4_vp2_sync_mbs1_mbc8_no_ckpt"])
    so = {}
    if "vp2" not in base:
        pp = rng.choice([1, 2, 4])
        so["pp_size"] = pp
        so["micro_batch_num"] = rng.choice([pp, 2 * pp, 8])
        if moe:
            so["ep_size"] = rng.choice([2, 4, 8])
        else:
            so["tp_size"] = rng.choice([1, 2, 4])
        so["world_size"] = rng.choice([8, 16])
    so["pp_comm_async"] = rng.random() < 0.5
    if rng.random() < 0.3:
        so["recompute_layer_num"] = rng.choice([1, 2])
    if rng.random() < 0.2:
        so["seq_len"] = rng.choice([2048, 8192])
    r = rng.random()
    if r < 0.25 and "vp2" not in base:
        so["enable_recompute"] = True
        so["recompute_granularity"] = rng.choice(["full_block", "attn_only", "mlp_only", "sdp_only", "selective_recompute"])
        so["recompute_layer_num"] = rng.choice([1, 2])
        if so["recompute_granularity"] == "selective_recompute":
            for flag in ("attn_recompute", "mlp_recompute", "mlp_rms_recompute"):
                so[flag] = rng.random() < 0.5
    elif r < 0.35 and "vp2" not in base:
        so["enable_recompute"] = False
    if not moe and so.get("tp_size", 1) > 1 and rng.random() < 0.3:
        so["enable_sequence_parallel"] = False
    if not moe and "vp2" not in base and rng.random() < 0.15:
        so.update(cp_size=2, seq_len=8192, cp_a2a_mode=rng.choice(["sync_cp", "async_cp"]))
    if moe and rng.random() < 0.5:
        so["dispatch_probs"] = True
    if rng.random() < 0.2:
        so["micro_batch_size"] = 2
    layers = rng.choice([4, 8]) if "vp2" not in base else 8
    return (base, model, system, layers, dict(MOE) if moe else {}, so, True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    stats = {"identical": 0, "both_reject": 0, "diff": 0}
    for i in range(n):
        case = sample(rng)
        res = []
        for impl in ("ref", "new"):
            try:
                res.append(run(impl, case))
            except Exception as ex:
                res.append({"error": f"{type(ex).__name__}: {str(ex)[:120]}"})
        a, b = res
        if "error" in a and "error" in b:
            stats["both_reject"] += 1
        elif a == b:
            stats["identical"] += 1
        else:
            stats["diff"] += 1
            print("DIFF", i, json.dumps(case))
            print("   ref:", json.dumps(a)[:500])
            print("   new:", json.dumps(b)[:500])
    print(json.dumps(stats))
    return stats["diff"]


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
