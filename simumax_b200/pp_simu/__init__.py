from .utils import cal_FandB_multi_layers, perf_dualpp  # noqa: F401
