"""simumax_b200 -- a B200-native analytical performance / memory model for distributed LLM training,
with a hand-written sm_100a calibration + measurement kernel suite.

Public entry points mirror the reference (MooreThreads/SimuMax) so existing scripts port by changing
one import:

    from simumax_b200 import PerfLLM
    perf = PerfLLM(); perf.configure(strategy_config=..., model_config=..., system_config=...)
    perf.run_estimate(); perf.analysis()          # or simply perf.perf()
"""
__version__ = "0.1.0"

_LAZY = {
    "PerfLLM": ("simumax_b200.core.perf_llm", "PerfLLM"),
    "StrategyConfig": ("simumax_b200.core.config", "StrategyConfig"),
    "SystemConfig": ("simumax_b200.core.config", "SystemConfig"),
    "ModelConfig": ("simumax_b200.core.config", "ModelConfig"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(mod), attr)
    raise AttributeError(name)
