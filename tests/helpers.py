import io
import os
import warnings
from contextlib import redirect_stdout

import pytest

from simumax_b200.configs import get_simu_model_config, get_simu_strategy_config, get_simu_system_config
from simumax_b200.core.config import ModelConfig, StrategyConfig, SystemConfig
from simumax_b200.core.perf_llm import PerfLLM

MOE = {"moe_pad_expert_input_to_capacity": True, "capacity": 1}


def reference_system(name="b200_bf16_ceperm"):
    """Path of a system config shipped by the reference (needed to replay its published numbers)."""
    try:
        return get_simu_system_config(name)
    except KeyError:
        pytest.skip(f"reference system config {name}.json not available")


def build(strategy, model, system="b200_nominal", layers=None, model_over=None, run=True, **strategy_over):
    warnings.filterwarnings("ignore")
    d = StrategyConfig.read_json_file(get_simu_strategy_config(strategy))
    st = StrategyConfig.init_from_dict(d)
    for k, v in strategy_over.items():
        setattr(st, k, v)
    sys_path = system if os.path.sep in system else get_simu_system_config(system)
    p = PerfLLM()
    p._search_verbose = False
    with redirect_stdout(io.StringIO()):
        p.configure(strategy_config=st, model_config=ModelConfig.init_from_config_file(get_simu_model_config(model)),
                    system_config=SystemConfig.init_from_config_file(sys_path))
        if layers:
            p.model_config.layer_num = layers
        for k, v in (model_over or {}).items():
            setattr(p.model_config, k, v)
        p.model_config.padded_vocab_size = True
        p.model_config.make_vocab_size_divisible_by = 128
        if run:
            p.run_estimate()
    return p


def headline(p):
    with redirect_stdout(io.StringIO()):
        r = p.analysis(None, console_log=False)
    return r
