"""Top-level helpers (reference ``simumax/utils.py``): config registry look-ups."""
from ..configs import (get_simu_model_config, get_simu_strategy_config, get_simu_system_config, list_models,  # noqa: F401
                       list_strategies, list_systems)
