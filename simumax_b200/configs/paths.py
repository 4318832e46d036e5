from __future__ import annotations

import json
import os
import tempfile
from typing import List

from .registry import MODEL_ZOO, STRATEGY_PRESETS, model_dict, strategy_dict

_HERE = os.path.dirname(os.path.abspath(__file__))
SYSTEM_DIR = os.path.join(_HERE, "system")
_CACHE = os.environ.get("SIMUMAX_B200_CONFIG_CACHE", os.path.join(tempfile.gettempdir(), "simumax_b200_configs"))


def _write(kind: str, name: str, data: dict) -> str:
    d = os.path.join(_CACHE, kind)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, name + ".json")
    text = json.dumps(data, indent=4)
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w", encoding="utf-8") as fh:
            fh.write(text)
    return path


def get_simu_model_config(name: str) -> str:
    if name not in MODEL_ZOO:
        raise KeyError(f"unknown model '{name}', available: {sorted(MODEL_ZOO)}")
    return _write("models", name, model_dict(name))


def get_simu_strategy_config(name: str) -> str:
    if name not in STRATEGY_PRESETS:
        raise KeyError(f"unknown strategy '{name}', available: {sorted(STRATEGY_PRESETS)}")
    return _write("strategy", name, strategy_dict(name))


def list_systems() -> List[str]:
    return sorted(f[:-5] for f in os.listdir(SYSTEM_DIR) if f.endswith(".json")) if os.path.isdir(SYSTEM_DIR) else []


def get_simu_system_config(name: str) -> str:
    path = os.path.join(SYSTEM_DIR, name + ".json")
    if os.path.exists(path):
        return path
    for root in (os.environ.get("SIMUMAX_REFERENCE_CONFIGS", ""), "/root/reference/configs",
                 os.path.join(os.path.dirname(os.path.dirname(_HERE)), "baseline", "_ref", "configs")):
        cand = os.path.join(root, "system", name + ".json") if root else ""
        if cand and os.path.exists(cand):
            return cand
    raise KeyError(f"unknown system '{name}', available: {list_systems()}")


def list_models() -> List[str]:
    return sorted(MODEL_ZOO)


def list_strategies() -> List[str]:
    return sorted(STRATEGY_PRESETS)


def materialize(target_dir: str) -> str:
    """Write the whole registry as a ``configs/{models,strategy,system}`` tree (handy for editing by hand)."""
    import shutil
    for name in MODEL_ZOO:
        os.makedirs(os.path.join(target_dir, "models"), exist_ok=True)
        json.dump(model_dict(name), open(os.path.join(target_dir, "models", name + ".json"), "w"), indent=4)
    for name in STRATEGY_PRESETS:
        os.makedirs(os.path.join(target_dir, "strategy"), exist_ok=True)
        json.dump(strategy_dict(name), open(os.path.join(target_dir, "strategy", name + ".json"), "w"), indent=4)
    os.makedirs(os.path.join(target_dir, "system"), exist_ok=True)
    for name in list_systems():
        shutil.copy(os.path.join(SYSTEM_DIR, name + ".json"), os.path.join(target_dir, "system", name + ".json"))
    return target_dir
