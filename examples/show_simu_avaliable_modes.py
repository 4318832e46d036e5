"""List every built-in model, strategy preset and system config name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simumax_b200.utils import show_simu_models, show_simu_strategy, show_simu_system  # noqa: E402

if __name__ == "__main__":
    show_simu_models()
    show_simu_system()
    show_simu_strategy()
