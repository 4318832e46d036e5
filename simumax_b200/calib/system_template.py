"""System-config skeletons: the starting point that calibration fills with measured tables.

``b200_template()`` encodes nominal B200 / NVLink-5 constants (dense bf16 2250 TFLOPS, fp8 4500, HBM3e 8000 GB/s
nominal, NVLink 900 GB/s per direction) with conservative default efficiencies; ``a100_pcie``-style PCIe tiers are
available through ``pcie_template()``.  The JSON schema is the reference's (docs/system.md:77-148) so any file
written here loads in both frameworks.   (reference: simu_tools/efficency_test/combine_efficiency.py:37-336)
"""
from __future__ import annotations

import copy
import json
from typing import Dict, Optional

NET_OPS_RING = {
    "all_reduce": {"scale": 2, "offset": -1},
    "all_gather": {"scale": 1, "offset": -1},
    "reduce_scatter": {"scale": 1, "offset": -1},
    "p2p": {"scale": 1, "offset": 0},
    "all2all": {"scale": 1, "offset": -1},
}


def _net(gbps: float, eff: float, latency_us: float, op_overrides: Optional[Dict[str, dict]] = None) -> dict:
    ops = copy.deepcopy(NET_OPS_RING)
    for k, v in (op_overrides or {}).items():
        ops[k].update(v)
    return {"processor_usage": 0.0, "bandwidth": {"gbps": gbps, "efficient_factor": eff, "latency_us": latency_us},
            "op": ops}


def b200_template(sys_name: str = "b200_nominal", mem_gbs: float = 180, num_per_node: int = 8) -> dict:
    op = lambda tflops, eff: {"tflops": tflops, "efficient_factor": eff, "accurate_efficient_factor": {}}
    bw = lambda eff, lat=30.0: {"gbps": 8000, "efficient_factor": eff, "latency_us": lat}
    return {
        "sys_name": sys_name,
        "num_per_node": num_per_node,
        "accelerator": {
            "backend": "cuda",
            "mem_gbs": mem_gbs,
            "op": {"default": op(2250, 0.40), "matmul": op(2250, 0.55), "fp8_matmul": op(4500, 0.45), "mxfp8_matmul": op(4500, 0.22),
                   "group_matmul": op(2250, 0.40), "fp8_group_matmul": op(4500, 0.35),
                   "sdp_fwd": op(2250, 0.45), "sdp_bwd": op(2250, 0.45)},
            "bandwidth": {"default": bw(0.70), "permute_fwd": bw(0.50), "permute_bwd": bw(0.50), "ce": bw(0.32),
                          "ce_fusion": bw(0.60)},
            "mode": "roofline",
        },
        "networks": {
            "low_intra_node": _net(900, 0.70, 8.0),
            "high_intra_node": _net(900, 0.70, 8.0, {"all_reduce": {"efficient_factor": 0.74, "latency_us": 5.5, "fixed_latency_us": 22.0},
                                                       "all_gather": {"efficient_factor": 0.67, "latency_us": 9.0, "fixed_latency_us": 23.0},
                                                       "reduce_scatter": {"efficient_factor": 0.67, "latency_us": 8.0, "fixed_latency_us": 25.0},
                                                       "all2all": {"efficient_factor": 0.60, "latency_us": 7.3, "fixed_latency_us": 20.0},
                                                       "p2p": {"efficient_factor": 0.80, "latency_us": 16.0}}),
            "inter_node": _net(400, 0.80, 15.0),
        },
    }


def pcie_template(sys_name: str = "pcie_node", tflops: float = 312, hbm_gbps: float = 1935, mem_gbs: float = 80) -> dict:
    op = lambda eff: {"tflops": tflops, "efficient_factor": eff, "accurate_efficient_factor": {}}
    bw = lambda eff: {"gbps": hbm_gbps, "efficient_factor": eff, "latency_us": 20.0}
    nets = {"intra_with_pcie": True}
    for tier, g in (("intra_node_pcie_2x", 24), ("intra_node_pcie_4x", 20), ("intra_node_pcie_8x", 16)):
        nets[tier] = _net(g, 0.8, 10.0)
    nets["inter_node"] = _net(25, 0.8, 20.0)
    return {"sys_name": sys_name, "num_per_node": 8,
            "accelerator": {"backend": "cuda", "mem_gbs": mem_gbs,
                            "op": {"default": op(0.5), "matmul": op(0.7), "fp8_matmul": op(0.7), "group_matmul": op(0.5),
                                   "fp8_group_matmul": op(0.5), "sdp_fwd": op(0.5), "sdp_bwd": op(0.5)},
                            "bandwidth": {"default": bw(0.7), "permute_fwd": bw(0.5), "permute_bwd": bw(0.5),
                                          "ce": bw(0.3), "ce_fusion": bw(0.5)},
                            "mode": "roofline"},
            "networks": nets}


def write(config: dict, path: str) -> str:
    with open(path, "w", encoding="utf-8") as fh:
        json.dump(config, fh, indent=4)
    return path
