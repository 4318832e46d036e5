"""Discrete-event replay of a tiny 2-layer Llama on tp2: writes a Chrome trace (open in chrome://tracing or Perfetto)
and a memory snapshot (``simumax_memory_snapshot_v1`` JSON + a torch memory_viz compatible pickle)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simumax_b200 import PerfLLM  # noqa: E402
from simumax_b200.configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config  # noqa: E402


def main(out_dir=None):
    out_dir = out_dir or os.path.join(tempfile.gettempdir(), "simumax_b200_simulate_demo")
    perf = PerfLLM()
    perf.configure(strategy_config=get_simu_strategy_config("tp2_pp1_dp4_mbs1"),
                   model_config=get_simu_model_config("llama2-tiny"),
                   system_config=get_simu_system_config(os.environ.get("SIMU_SYSTEM", "b200_nominal")))
    perf.model_config.layer_num = 2
    perf.run_estimate()
    result = perf.simulate(out_dir, merge_lanes=True)
    print(f"trace end {result['end_time_ms']:.3f} ms; artifacts in {out_dir}:")
    for f in sorted(os.listdir(out_dir)):
        print("  ", f)
    return result


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
