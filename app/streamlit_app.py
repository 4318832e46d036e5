"""Interactive front-end over ``PerfLLM``:  ``streamlit run app/streamlit_app.py``.

Pick a system / model / strategy preset, tweak parallelism, batch, recompute and hardware numbers, then evaluate:
the page shows the headline metrics, per-stage memory, the time breakdown, efficiency-table hits and misses, warnings
from the strategy validator, and lets you download the full analysis (JSON zip) or a Chrome trace from ``simulate()``.

The logic lives in :class:`Session` (plain Python, unit-tested without streamlit); the ``render_*`` functions are the UI.
(reference: app/streamlit_app.py — sidebar quick config, hardware / parameter / recompute panels, result dashboard.)
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile
import warnings
import zipfile
from dataclasses import dataclass, field
from typing import Dict, List, Optional

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from simumax_b200 import PerfLLM  # noqa: E402
from simumax_b200.configs import (get_simu_model_config, get_simu_strategy_config, get_simu_system_config,  # noqa: E402
                                  list_models, list_strategies, list_systems)
from simumax_b200.core.config import ModelConfig, StrategyConfig, SystemConfig  # noqa: E402

QUICK_PRESETS = {
    "single node (8 GPU)": dict(world_size=8, tp_size=1, pp_size=1, ep_size=1, micro_batch_size=1, micro_batch_num=8, seq_len=4096),
    "4 nodes (32 GPU)": dict(world_size=32, tp_size=2, pp_size=2, ep_size=1, micro_batch_size=1, micro_batch_num=16, seq_len=4096),
    "cluster (256 GPU)": dict(world_size=256, tp_size=8, pp_size=4, ep_size=1, micro_batch_size=1, micro_batch_num=32, seq_len=8192),
}
RECOMPUTE_KINDS = ["none", "full_block", "attn_only", "mlp_only", "sdp_only", "selective_recompute"]


@dataclass
class Evaluation:
    ok: bool
    cost: dict = field(default_factory=dict)
    memory: dict = field(default_factory=dict)
    warnings: List[str] = field(default_factory=list)
    error: Optional[str] = None
    hits: Dict[str, int] = field(default_factory=dict)
    misses: Dict[str, List[str]] = field(default_factory=dict)
    log: str = ""


class Session:
    """Everything the page does, without any UI dependency."""

    def __init__(self, system: str, model: str, strategy: str):
        self.system_dict = json.load(open(get_simu_system_config(system)))
        self.model_dict = json.load(open(get_simu_model_config(model)))
        self.strategy_dict = json.load(open(get_simu_strategy_config(strategy)))
        self.perf: Optional[PerfLLM] = None
        self.runtime_semantics = False           # price simumax_b200.runtime (fused kernels) instead of Megatron-LM + NCCL

    def apply(self, strategy: Optional[dict] = None, model: Optional[dict] = None, hardware: Optional[dict] = None):
        self.strategy_dict.update(strategy or {})
        self.model_dict.update(model or {})
        hw = hardware or {}
        acc = self.system_dict["accelerator"]
        if "mem_gbs" in hw:
            acc["mem_gbs"] = hw["mem_gbs"]
        if "tflops" in hw:
            for op in acc["op"].values():
                op["tflops"] = hw["tflops"]
        if "hbm_gbps" in hw:
            for bw in acc["bandwidth"].values():
                bw["gbps"] = hw["hbm_gbps"]
        for net in ("high_intra_node", "inter_node"):
            if f"{net}_gbps" in hw and net in self.system_dict["networks"]:
                self.system_dict["networks"][net]["bandwidth"]["gbps"] = hw[f"{net}_gbps"]

    def set_recompute(self, kind: str, layers: int = 0):
        s = self.strategy_dict
        if kind == "none":
            s.update(enable_recompute=False, recompute_granularity=None, recompute_layer_num=0)
        else:
            s.update(enable_recompute=True, recompute_granularity=kind, recompute_layer_num=layers)

    def evaluate(self) -> Evaluation:
        buf = io.StringIO()
        caught: List[str] = []
        try:
            with contextlib.redirect_stdout(buf), warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                perf = PerfLLM()
                strategy = StrategyConfig.init_from_dict(dict(self.strategy_dict))
                if self.runtime_semantics:
                    strategy.apply_runtime_semantics()
                perf.configure(strategy_config=strategy,
                               model_config=ModelConfig.init_from_dict(dict(self.model_dict)),
                               system_config=SystemConfig.init_from_dict(json.loads(json.dumps(self.system_dict))))
                perf.run_estimate()
                cost = perf.analysis_cost().data
                mem = perf.analysis_mem().data
                caught = [str(x.message) for x in w]
            self.perf = perf
            return Evaluation(True, cost, mem, caught, None, {k: len(v) for k, v in perf.system.hit_efficiency.items()},
                              {k: list(v) for k, v in perf.system.miss_efficiency.items()}, buf.getvalue())
        except (AssertionError, ValueError, KeyError, NotImplementedError) as e:
            return Evaluation(False, warnings=caught, error=f"{type(e).__name__}: {e}", log=buf.getvalue())

    def model_summary(self) -> Dict[str, object]:
        """Architecture facts shown in the "model details" panel (parameter counts, FLOPs per token, MoE / MLA flags)."""
        m = ModelConfig.init_from_dict(dict(self.model_dict))
        out = {"model_name": m.model_name, "layers": m.layer_num, "hidden": m.hidden_size, "heads": m.head_num,
               "kv_heads": m.kv_head_num, "ffn": m.intermediate_size, "vocab": m.vocab_size, "experts": m.expert_num,
               "attention": getattr(m, "attention_type", None) or "mha/gqa"}
        for name in ("param_numel", "activated_param_numel", "flops_per_token"):
            attr = getattr(m, name, None)
            try:
                out[name] = attr() if callable(attr) else attr
            except Exception:                    # some counters need a strategy (sequence length); they are optional here
                pass
        return out

    def suggestions(self, ev: Evaluation) -> List[str]:
        """Plain-language hints derived from one evaluation (reference app: "提示/建议" panel)."""
        from simumax_b200.core.utils import parse_bytes
        tips: List[str] = []
        if not ev.ok:
            return [f"configuration rejected: {ev.error}"]
        s = self.strategy_dict
        budget = float(self.system_dict["accelerator"]["mem_gbs"]) * 2 ** 30
        stages = ev.memory if "peak_mem" not in ev.memory else {"stage": ev.memory}
        peaks = {k: parse_bytes(v["peak_mem"]) for k, v in stages.items() if isinstance(v, dict) and "peak_mem" in v}
        worst = max(peaks.values()) if peaks else 0
        if worst > budget:
            over = (worst - budget) / 2 ** 30
            tips.append(f"peak memory exceeds the {budget / 2 ** 30:.0f} GiB device by {over:.1f} GiB: raise tp / pp, enable "
                        f"recompute (full_block on a few layers first) or shrink the micro-batch")
        elif worst and worst < 0.5 * budget and s.get("enable_recompute"):
            tips.append("less than half of the device memory is used with recompute on: recompute fewer layers or none")
        elif worst and worst < 0.5 * budget:
            tips.append("less than half of the device memory is used: a larger micro-batch or less model parallelism would raise MFU")
        br = ev.cost.get("breakdown_result", {})
        ms = lambda v: float(str(v).split()[0]) if v is not None else 0.0
        total = ms(ev.cost.get("duration_time_per_iter"))
        if total > 0:
            if ms(br.get("intra_exposed_time")) > 0.15 * total:
                tips.append("tensor / expert-parallel collectives take more than 15 % of the iteration: lower tp (or ep), or turn on "
                            "sequence parallelism; on this repo's runtime set tp_overlap_model (fused AG->GEMM / GEMM->RS kernels)")
            if ms(br.get("dp_exposed_time")) > 0.10 * total:
                tips.append("data-parallel gradient / parameter collectives exceed 10 % of the iteration: more micro-batches per step "
                            "amortise them (dp_overlap_model describes the overlapped runtime)")
            stage_t = ev.cost.get("all_chunk_times", {})
            bubbles = [ms(v.get("bubble_time")) for v in stage_t.values() if isinstance(v, dict)]
            if bubbles and max(bubbles) > 0.15 * total:
                tips.append("pipeline bubble above 15 %: increase micro_batch_num or use interleaved (virtual) stages")
        if ev.misses:
            n = sum(len(v) for v in ev.misses.values())
            tips.append(f"{n} operator shapes are not in the efficiency table of this system config and fall back to defaults: "
                        f"run `python -m simumax_b200.calib.calibrate ops --supplement <system.json>` on the target machine")
        return tips

    def report_zip(self, ev: Evaluation) -> bytes:
        out = io.BytesIO()
        with zipfile.ZipFile(out, "w", zipfile.ZIP_DEFLATED) as z:
            z.writestr("strategy.json", json.dumps(self.strategy_dict, indent=2))
            z.writestr("model.json", json.dumps(self.model_dict, indent=2))
            z.writestr("system.json", json.dumps(self.system_dict, indent=2))
            z.writestr("cost.json", json.dumps(ev.cost, indent=2, default=str))
            z.writestr("memory.json", json.dumps(ev.memory, indent=2, default=str))
            z.writestr("efficiency_misses.json", json.dumps(ev.misses, indent=2))
        return out.getvalue()

    def trace_json(self) -> bytes:
        assert self.perf is not None, "evaluate first"
        d = tempfile.mkdtemp(prefix="smx_app_")
        with contextlib.redirect_stdout(io.StringIO()):
            self.perf.simulate(d, merge_lanes=True)
        return open(os.path.join(d, "tracing_logs.json"), "rb").read()


def headline(ev: Evaluation) -> Dict[str, str]:
    c = ev.cost
    peaks = ev.memory if "peak_mem" in ev.memory else {k: v for k, v in ev.memory.items() if isinstance(v, dict)}
    peak = peaks.get("peak_mem") if "peak_mem" in peaks else ", ".join(f"{k}: {v.get('peak_mem')}" for k, v in peaks.items())
    return {"iteration time": str(c.get("duration_time_per_iter")), "tokens/s/GPU": f"{c.get('throughput_per_accelerator', 0):,.1f}",
            "MFU": f"{100 * c.get('mfu', 0):.2f} %", "peak memory": str(peak)}


# ------------------------------------------------------------------------------------------------------------- UI
def main():  # pragma: no cover - needs streamlit
    import streamlit as st
    st.set_page_config(page_title="simumax_b200", layout="wide", initial_sidebar_state="expanded")
    st.title("simumax_b200 — LLM training performance model")
    with st.sidebar:
        st.header("Quick config")
        system = st.selectbox("system", list_systems(), index=0)
        model = st.selectbox("model", list_models(), index=list_models().index("llama3-8b"))
        strategy = st.selectbox("strategy preset", list_strategies())
        quick = st.selectbox("scale preset", ["(keep strategy preset)"] + list(QUICK_PRESETS))
    sess = Session(system, model, strategy)
    if quick in QUICK_PRESETS:
        sess.apply(strategy=QUICK_PRESETS[quick])
    s = sess.strategy_dict
    tab_par, tab_hw, tab_rc = st.tabs(["Parallelism & batch", "Hardware", "Recompute"])
    with tab_par:
        cols = st.columns(4)
        upd = {}
        for i, key in enumerate(["world_size", "tp_size", "pp_size", "ep_size", "etp_size", "cp_size", "seq_len",
                                 "micro_batch_size", "micro_batch_num", "zero_state"]):
            upd[key] = cols[i % 4].number_input(key, min_value=0 if key == "zero_state" else 1, value=int(s.get(key, 1)), step=1)
        upd["enable_sequence_parallel"] = st.checkbox("sequence parallel", value=bool(s.get("enable_sequence_parallel", True)))
        upd["fp8"] = st.checkbox("fp8 linears", value=bool(s.get("fp8", False)))
        upd["fp8_recipe"] = st.selectbox("fp8 recipe", ["tensor", "mxfp8"], index=0)
        sess.runtime_semantics = st.checkbox("price this repo's fused runtime (instead of Megatron-LM + NCCL)", value=False)
        layers = st.number_input("layer_num", min_value=1, value=int(sess.model_dict.get("layer_num", 32)))
        with st.expander("uneven pipeline split (advanced)"):
            first = st.number_input("num_layers_in_first_pipeline_stage (0 = even)", min_value=0, value=0)
            last = st.number_input("num_layers_in_last_pipeline_stage (0 = even)", min_value=0, value=0)
            upd["num_layers_in_first_pipeline_stage"] = int(first) or None
            upd["num_layers_in_last_pipeline_stage"] = int(last) or None
        with st.expander("model parameters"):
            mupd = {}
            for key in ("hidden_size", "head_num", "kv_head_num", "intermediate_size", "vocab_size", "expert_num", "topk"):
                if key in sess.model_dict and isinstance(sess.model_dict[key], int):
                    mupd[key] = int(st.number_input(key, min_value=1, value=int(sess.model_dict[key])))
            sess.apply(model=mupd)
        sess.apply(strategy=upd, model={"layer_num": int(layers)})
        with st.expander("model details", expanded=False):
            st.json(sess.model_summary())
    with tab_hw:
        acc = sess.system_dict["accelerator"]
        c1, c2, c3 = st.columns(3)
        hw = {"mem_gbs": c1.number_input("HBM (GiB)", value=float(acc["mem_gbs"])),
              "tflops": c2.number_input("dense bf16 TFLOPS", value=float(acc["op"]["matmul"]["tflops"])),
              "hbm_gbps": c3.number_input("HBM GB/s", value=float(acc["bandwidth"]["default"]["gbps"])),
              "high_intra_node_gbps": c1.number_input("NVLink GB/s/dir", value=float(sess.system_dict["networks"]["high_intra_node"]["bandwidth"]["gbps"])),
              "inter_node_gbps": c2.number_input("inter-node GB/s", value=float(sess.system_dict["networks"]["inter_node"]["bandwidth"]["gbps"]))}
        sess.apply(hardware=hw)
    with tab_rc:
        kind = st.selectbox("granularity", RECOMPUTE_KINDS)
        n = st.number_input("recompute_layer_num", min_value=0, value=0)
        sess.set_recompute(kind, int(n))
    if st.button("Evaluate", use_container_width=True):
        ev = sess.evaluate()
        if not ev.ok:
            st.error(ev.error)
        else:
            for col, (k, v) in zip(st.columns(4), headline(ev).items()):
                col.metric(k, v)
            left, right = st.columns(2)
            left.subheader("Memory"); left.json(ev.memory, expanded=False)
            right.subheader("Time breakdown"); right.json(ev.cost.get("breakdown_result", ev.cost), expanded=False)
            tips = sess.suggestions(ev)
            if tips:
                st.subheader("Suggestions")
                for t in tips:
                    st.info(t)
            with st.expander("achieved collective bandwidths"):
                st.json(json.loads(json.dumps(sess.perf.system.real_comm_bw, default=str)))
            st.subheader("Efficiency table")
            st.write({"hits": ev.hits, "misses": {k: len(v) for k, v in ev.misses.items()}})
            st.download_button("Download analysis (zip)", sess.report_zip(ev), file_name="simumax_b200_analysis.zip")
            st.download_button("Download Chrome trace (simulate)", sess.trace_json(), file_name="tracing_logs.json")
        for wmsg in ev.warnings:
            st.warning(wmsg)


if __name__ == "__main__":
    main()
