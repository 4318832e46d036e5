"""Release charts without a plotting dependency (pure-Python SVG):
  * long-context CP-a2a sweep: predicted iteration time / peak memory of Llama-3-70B(L12) tp2 x cp{1,2,4} at seq 8k..64k,
  * predicted-vs-measured bars from a perf_vs_real JSON (tools/b200/perf_vs_real.py).
    python tools/b200/plot_release_charts.py --out profiles/charts [--pvr profiles/perf_vs_real_1gpu.json]
(reference: tools/b200/plot_release_charts.py draws the CP A2A bar chart with matplotlib.)
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def bar_chart_svg(title, groups, series, values, unit, width=900, height=420):
    """groups: x labels; series: legend names; values[s][g] numbers. Returns SVG text."""
    colors = ["#3b6fb6", "#d9822b", "#5aa469", "#a04a9a", "#777777"]
    pad_l, pad_b, pad_t = 70, 60, 50
    W, H = width - pad_l - 20, height - pad_b - pad_t
    vmax = max(max(v) for v in values) * 1.1 or 1.0
    gw = W / len(groups)
    bw = gw * 0.8 / len(series)
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{width}" height="{height}" font-family="sans-serif" font-size="12">',
           f'<rect width="100%" height="100%" fill="white"/><text x="{width / 2}" y="24" text-anchor="middle" font-size="16">{title}</text>']
    for i in range(6):
        y = pad_t + H - H * i / 5
        out.append(f'<line x1="{pad_l}" x2="{pad_l + W}" y1="{y}" y2="{y}" stroke="#ddd"/>'
                   f'<text x="{pad_l - 6}" y="{y + 4}" text-anchor="end">{vmax * i / 5:.3g}</text>')
    out.append(f'<text x="16" y="{pad_t + H / 2}" transform="rotate(-90 16 {pad_t + H / 2})" text-anchor="middle">{unit}</text>')
    for g, label in enumerate(groups):
        x0 = pad_l + g * gw + gw * 0.1
        for s in range(len(series)):
            v = values[s][g]
            h = H * v / vmax
            out.append(f'<rect x="{x0 + s * bw:.1f}" y="{pad_t + H - h:.1f}" width="{bw * 0.92:.1f}" height="{h:.1f}" fill="{colors[s % 5]}"/>'
                       f'<text x="{x0 + s * bw + bw / 2:.1f}" y="{pad_t + H - h - 3:.1f}" text-anchor="middle" font-size="10">{v:.3g}</text>')
        out.append(f'<text x="{pad_l + g * gw + gw / 2:.1f}" y="{pad_t + H + 18}" text-anchor="middle">{label}</text>')
    for s, name in enumerate(series):
        out.append(f'<rect x="{pad_l + s * 170}" y="{height - 22}" width="12" height="12" fill="{colors[s % 5]}"/>'
                   f'<text x="{pad_l + s * 170 + 18}" y="{height - 12}">{name}</text>')
    out.append("</svg>")
    return "\n".join(out)


def cp_sweep(system):
    from simumax_b200 import PerfLLM
    from simumax_b200.configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config
    from simumax_b200.core.config import StrategyConfig
    from simumax_b200.core.utils import parse_bytes, parse_time_ms
    seqs, cps = [8192, 16384, 32768, 65536], [1, 2, 4]
    t = [[0.0] * len(seqs) for _ in cps]
    m = [[0.0] * len(seqs) for _ in cps]
    for ci, cp in enumerate(cps):
        for si, seq in enumerate(seqs):
            st = StrategyConfig.init_from_config_file(get_simu_strategy_config("tp2_pp1_dp4_mbs1"))
            st.cp_size, st.seq_len, st.micro_batch_num = cp, seq, 4
            if cp > 1:
                st.cp_comm_type = "a2a"
            p = PerfLLM()
            with contextlib.redirect_stdout(io.StringIO()):
                p.configure(strategy_config=st, model_config=get_simu_model_config("llama3-70b"), system_config=get_simu_system_config(system))
                p.model_config.layer_num = 12
                p.model_config.padded_vocab_size, p.model_config.make_vocab_size_divisible_by = True, 128
                p.run_estimate()
                t[ci][si] = parse_time_ms(p.analysis_cost().data["duration_time_per_iter"])
                m[ci][si] = parse_bytes(p.analysis_mem().data["peak_mem"]) / 2 ** 30
    return seqs, cps, t, m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="profiles/charts")
    ap.add_argument("--system", default="b200_nominal")
    ap.add_argument("--pvr", default=None)
    ap.add_argument("--bench-glob", default=None, help="e.g. 'profiles/bench_r2_*n?.json'")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    seqs, cps, t, m = cp_sweep(a.system)
    open(os.path.join(a.out, "cp_a2a_time.svg"), "w").write(
        bar_chart_svg(f"Llama-3-70B L12 tp2 x cp (a2a), {a.system}: iteration time", [f"seq {s}" for s in seqs], [f"cp{c}" for c in cps], t, "ms / iteration"))
    open(os.path.join(a.out, "cp_a2a_mem.svg"), "w").write(
        bar_chart_svg(f"Llama-3-70B L12 tp2 x cp (a2a), {a.system}: peak memory", [f"seq {s}" for s in seqs], [f"cp{c}" for c in cps], m, "GiB"))
    if a.pvr and os.path.exists(a.pvr):
        rows = json.load(open(a.pvr))["rows"]
        labels = [str(i + 1) for i in range(len(rows))]
        wide = max(900, 60 * len(rows))
        open(os.path.join(a.out, "perf_vs_real_time.svg"), "w").write(
            bar_chart_svg("iteration time: predicted vs measured (row numbers as in perf_vs_real_round2.md)", labels, ["predicted", "measured"],
                          [[r["pred_ms"] for r in rows], [r["real_ms"] for r in rows]], "ms", width=wide))
        open(os.path.join(a.out, "perf_vs_real_mem.svg"), "w").write(
            bar_chart_svg("peak memory: predicted vs measured", labels, ["predicted", "measured"],
                          [[r["pred_gib"] for r in rows], [r["real_gib"] for r in rows]], "GiB", width=wide))
        open(os.path.join(a.out, "perf_vs_real_abs_err.svg"), "w").write(
            bar_chart_svg("|error| of the prediction", labels, ["time", "memory"],
                          [[abs(r["time_err_pct"]) for r in rows], [abs(r["mem_err_pct"]) for r in rows]], "%", width=wide))
        with open(os.path.join(a.out, "perf_vs_real_rows.txt"), "w") as f:
            for i, r in enumerate(rows):
                f.write(f"{i + 1:2d}  {r['case']}  ({r['world']} GPU, {r['parallelism']})\n")
    if a.bench_glob:
        import glob
        own, ref = {}, {}
        for fn in sorted(glob.glob(a.bench_glob)):
            for line in open(fn):
                if line.startswith("{"):
                    j = json.loads(line)
                    (ref if j.get("impl") == "reference" else own)[j["n_gpus"]] = j
        ns = sorted(set(own) & set(ref)) or sorted(own)
        if ns:
            series, vals = ["this repo (device)", "this repo (end to end)"], [[own[n]["value"] for n in ns], [own[n]["e2e"]["value"] for n in ns]]
            if all(n in ref for n in ns):
                series.append("library arm (end to end)")
                vals.append([ref[n]["e2e"]["value"] for n in ns])
            open(os.path.join(a.out, "bench_scaling.svg"), "w").write(
                bar_chart_svg("Llama-3-8B seq 4096, tp = N, 8 micro-batches: training tokens/s", [f"N={n}" for n in ns], series, vals, "tokens/s"))
    print("wrote", sorted(os.listdir(a.out)))


if __name__ == "__main__":
    main()
