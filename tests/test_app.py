"""The Streamlit front-end's logic layer (Session) runs without streamlit."""
import importlib.util
import json
import os
import zipfile
import io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("streamlit_app", os.path.join(ROOT, "app", "streamlit_app.py"))
    import sys
    mod = importlib.util.module_from_spec(spec)
    sys.modules["streamlit_app"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_session_evaluate_and_report():
    app = _load()
    s = app.Session("b200_nominal", "llama3-8b", "tp2_pp1_dp4_mbs1")
    s.apply(strategy={"micro_batch_num": 4}, hardware={"mem_gbs": 180, "tflops": 2000})
    s.set_recompute("full_block", 4)
    ev = s.evaluate()
    assert ev.ok and ev.cost["throughput_per_accelerator"] > 0
    head = app.headline(ev)
    assert set(head) == {"iteration time", "tokens/s/GPU", "MFU", "peak memory"}
    with zipfile.ZipFile(io.BytesIO(s.report_zip(ev))) as z:
        assert {"strategy.json", "model.json", "system.json", "cost.json", "memory.json"} <= set(z.namelist())
        assert json.loads(z.read("strategy.json"))["recompute_layer_num"] == 4
    trace = json.loads(s.trace_json())
    assert len(trace["traceEvents"] if isinstance(trace, dict) else trace) > 100
    # slower hardware -> slower iteration
    fast = ev.cost["throughput_per_accelerator"]
    s.apply(hardware={"tflops": 1000})
    assert s.evaluate().cost["throughput_per_accelerator"] < fast


def test_session_reports_invalid_strategy():
    app = _load()
    s = app.Session("b200_nominal", "llama3-8b", "tp1_pp1_dp8_mbs1")
    s.apply(strategy={"tp_size": 3})
    ev = s.evaluate()
    assert not ev.ok and "divisible" in ev.error


def test_model_summary_and_suggestions():
    app = _load()
    s = app.Session("b200_nominal", "llama3-70b", "tp1_pp1_dp8_mbs1")
    info = s.model_summary()
    assert info["layers"] == 80 and info["hidden"] == 8192 and info["kv_heads"] == 8
    ev = s.evaluate()                      # 70B on tp1 / pp1 cannot fit on one 180 GB device
    tips = s.suggestions(ev)
    assert ev.ok and any("exceeds" in t for t in tips), tips
    s2 = app.Session("b200_nominal", "llama3-8b", "tp8_pp1_dp1_mbs1")
    s2.apply(strategy={"micro_batch_num": 1})
    tips2 = s2.suggestions(s2.evaluate())
    assert any("collectives take more than 15" in t for t in tips2) and any("less than half" in t for t in tips2), tips2
    bad = app.Session("b200_nominal", "llama3-8b", "tp1_pp1_dp8_mbs1")
    bad.apply(strategy={"tp_size": 3})
    assert bad.suggestions(bad.evaluate())[0].startswith("configuration rejected")


def test_session_runtime_semantics_toggle():
    app = _load()
    s = app.Session("b200_sm100a_calibrated", "llama3-8b", "tp8_pp1_dp1_mbs1")
    meg = s.evaluate()
    s.runtime_semantics = True
    fused = s.evaluate()
    assert meg.ok and fused.ok
    assert fused.cost["throughput_per_accelerator"] > 1.15 * meg.cost["throughput_per_accelerator"]
