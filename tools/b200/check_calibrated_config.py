"""Replay the calibration workloads on a system config and report efficiency-table coverage + headline predictions."""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from simumax_b200 import PerfLLM  # noqa: E402
from simumax_b200.calib.shapes import DEFAULT_WORKLOADS  # noqa: E402
from simumax_b200.configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config  # noqa: E402


def main(system):
    print(f"{'model':12s} {'strategy':24s} {'iter':>12s} {'tok/s/gpu':>10s} {'mfu':>6s} {'peak':>12s}  hits/misses")
    for model, strategy, m_over, _ in DEFAULT_WORKLOADS:
        p = PerfLLM()
        with contextlib.redirect_stdout(io.StringIO()):
            p.configure(strategy_config=get_simu_strategy_config(strategy), model_config=get_simu_model_config(model),
                        system_config=get_simu_system_config(system))
            for k, v in m_over.items():
                setattr(p.model_config, k, v)
            p.model_config.padded_vocab_size, p.model_config.make_vocab_size_divisible_by = True, 128
            p.run_estimate()
            cost, mem = p.analysis_cost().data, p.analysis_mem().data
        peak = mem.get("peak_mem") or "/".join(str(v.get("peak_mem")) for v in mem.values() if isinstance(v, dict))
        hits = sum(len(v) for v in p.system.hit_efficiency.values())
        miss = sum(len(v) for v in p.system.miss_efficiency.values())
        print(f"{model:12s} {strategy:24s} {str(cost['duration_time_per_iter']):>12s} {cost['throughput_per_accelerator']:10.1f} "
              f"{cost['mfu']:6.3f} {peak:>12s}  {hits}/{miss}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "b200_sm100a_calibrated")
