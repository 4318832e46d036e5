"""Config registry: name -> JSON path for the built-in model zoo, strategy presets and system configs."""
from .paths import (get_simu_model_config, get_simu_strategy_config, get_simu_system_config, list_models,  # noqa: F401
                    list_strategies, list_systems, materialize)
