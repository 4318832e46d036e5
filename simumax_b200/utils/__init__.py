"""Top-level helpers (reference ``simumax/utils.py``): config registry look-ups."""
from ..configs import (get_simu_model_config, get_simu_strategy_config, get_simu_system_config, list_models,  # noqa: F401
                       list_strategies, list_systems)


def _show(title, names):
    print(f"{title} ({len(names)}):")
    for n in names:
        print(f"  - {n}")


def show_simu_models():
    _show("models", list_models())


def show_simu_strategy():
    _show("strategy presets", list_strategies())


def show_simu_system():
    _show("system configs", list_systems())
