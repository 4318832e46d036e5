"""DualPipe (DeepSeek-V3 bidirectional pipeline) closed forms.

DualPipe overlaps one micro-batch's forward (F) with another's backward (B) inside each stage and splits the
backward into input-grad (B) and weight-grad (W).  With PP ranks, per-chunk times F, B (full backward), W and the
overlapped F&B time, the critical path of the schedule is (DeepSeek-V3 tech report, table 2):

    bubble   = (PP/2 - 1) * (F&B + B - 3W)
    duration = bubble + mbc * (F + B)  (per direction pair, two chunks per rank)

These are stand-alone helpers (not wired into PerfLLM), mirroring reference pp_simu/utils.py.
"""
from __future__ import annotations

from typing import Dict, List, Sequence


def cal_FandB_multi_layers(f_times: Sequence[float], b_times: Sequence[float], comm_f: Sequence[float] = None,
                           comm_b: Sequence[float] = None) -> Dict[str, float]:
    """Overlapped forward&backward time of a chunk of layers when F compute hides B communication and vice versa.

    Per layer the overlapped pair costs max(F_compute + B_compute, F_comm + B_comm) (compute and all-to-all run on
    different resources); returns the serial time, the overlapped time and the saving.
    """
    n = len(f_times)
    comm_f = list(comm_f) if comm_f is not None else [0.0] * n
    comm_b = list(comm_b) if comm_b is not None else [0.0] * n
    serial = sum(f_times) + sum(b_times) + sum(comm_f) + sum(comm_b)
    overlapped = sum(max(f + b, cf + cb) for f, b, cf, cb in zip(f_times, b_times, comm_f, comm_b))
    return {"serial": serial, "overlapped": overlapped, "saved": serial - overlapped}


def perf_dualpp(pp_size: int, micro_batch_num: int, f: float, b: float, w: float, f_and_b: float = None,
                flops_per_microbatch: float = None, peak_tflops: float = None) -> Dict[str, float]:
    """Iteration time / bubble / (optional) MFU of DualPipe vs 1F1B and ZB1P for the same chunk times.

    ``f``: forward of one chunk, ``b``: full backward of one chunk, ``w``: weight-grad part of the backward,
    ``f_and_b``: time of a mutually overlapped forward+backward pair (defaults to f + b).
    """
    assert pp_size % 2 == 0, "DualPipe needs an even number of pipeline ranks"
    fb = f + b if f_and_b is None else f_and_b
    bubble_1f1b = (pp_size - 1) * (f + b)
    bubble_zb1p = (pp_size - 1) * (f + b - 2 * w)
    bubble_dual = (pp_size / 2 - 1) * (fb + b - 3 * w)
    work = micro_batch_num * (f + b)
    out = {"bubble_1f1b": bubble_1f1b, "bubble_zb1p": bubble_zb1p, "bubble_dualpipe": bubble_dual,
           "duration_1f1b": work + bubble_1f1b, "duration_zb1p": work + bubble_zb1p,
           "duration_dualpipe": work + bubble_dual,
           "activation_multiplier_1f1b": pp_size, "activation_multiplier_dualpipe": pp_size + 1,
           "param_multiplier_dualpipe": 2}
    if flops_per_microbatch and peak_tflops:
        for k in ("1f1b", "zb1p", "dualpipe"):
            out[f"mfu_{k}"] = flops_per_microbatch * micro_batch_num / (out[f"duration_{k}"] / 1e3) / 1e12 / peak_tflops
    return out


def plot_FandB_timeline(f_times: List[float], b_times: List[float], path: str = "fandb_timeline.png"):
    """Two-row Gantt chart of the per-layer F / B overlap (requires matplotlib)."""
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    fig, ax = plt.subplots(figsize=(10, 2.4))
    t = 0.0
    for i, (f, b) in enumerate(zip(f_times, b_times)):
        ax.broken_barh([(t, f)], (10, 8), facecolors="tab:blue")
        ax.broken_barh([(t, b)], (0, 8), facecolors="tab:orange")
        t += max(f, b)
    ax.set_yticks([4, 14]); ax.set_yticklabels(["B", "F"]); ax.set_xlabel("ms")
    fig.tight_layout(); fig.savefig(path); plt.close(fig)
    return path
