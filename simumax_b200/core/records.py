"""Additive numeric records shared by the op IR, the stage chunks and the orchestrator.

Counterparts of the reference's ``ModuleComputeInfo / ActivationInfo / ModuleMemoryInfo /
ModuleCostInfo / PeakPoint`` (simumax/core/model_struct.py:41-412, transformer/language_model.py:12-97).
Everything is kept as raw numbers; human formatting happens only at the JSON edge (utils.fmt).
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field, fields
from typing import Optional, Set, Tuple


class RecomputeStatus:
    NO_RECOMPUTE = "no_recompute"
    FIRST = "first"
    MIDDLE = "middle"
    LAST = "last"


_FIELD_NAMES: dict = {}


def _add_dataclass(a, b):
    cls = type(a)
    names = _FIELD_NAMES.get(cls)
    if names is None:                      # dataclasses.fields() walks the class on every call: cache the names per record type
        names = _FIELD_NAMES[cls] = tuple(f.name for f in fields(cls))
    da, db = a.__dict__, b.__dict__
    return cls(**{n: da[n] + db[n] for n in names})


@dataclass
class ComputeInfo:
    fwd_flops: float = 0
    recompute_flops: float = 0
    bwd_grad_w_flops: float = 0
    bwd_grad_act_flops: float = 0
    fwd_accessed_mem: float = 0
    recompute_accessed_mem: float = 0
    bwd_grad_w_accessed_mem: float = 0
    bwd_grad_act_accessed_mem: float = 0

    @property
    def bwd_flops(self):
        return self.bwd_grad_w_flops + self.bwd_grad_act_flops

    @property
    def bwd_accessed_mem(self):
        return self.bwd_grad_w_accessed_mem + self.bwd_grad_act_accessed_mem

    def get_all_flops(self):
        return [self.fwd_flops, self.bwd_grad_act_flops, self.bwd_grad_w_flops]

    def get_all_accessed_mem(self):
        return [self.fwd_accessed_mem, self.bwd_grad_act_accessed_mem, self.bwd_grad_w_accessed_mem]

    def __add__(self, other: "ComputeInfo") -> "ComputeInfo":
        return _add_dataclass(self, other)


@dataclass
class CostInfo:
    """All times in milliseconds."""
    fwd_compute_time: float = 0
    recompute_compute_time: float = 0
    bwd_grad_w_time: float = 0
    bwd_grad_act_time: float = 0
    fwd_net_time: float = 0
    recompute_net_time: float = 0
    bwd_grad_w_net_time: float = 0
    bwd_grad_act_net_time: float = 0
    fwd_net_exposed_time: float = 0
    recompute_net_exposed_time: float = 0
    bwd_net_exposed_time: float = 0

    @property
    def fwd_time(self):
        return self.fwd_compute_time + self.fwd_net_exposed_time

    @property
    def recompute_time(self):
        return self.recompute_compute_time + self.recompute_net_exposed_time

    @property
    def bwd_compute_time(self):
        return self.bwd_grad_w_time + self.bwd_grad_act_time

    @property
    def bwd_time(self):
        return self.bwd_grad_w_time + self.bwd_grad_act_time + self.bwd_net_exposed_time

    @property
    def bwd_net_time(self):
        return self.bwd_grad_w_net_time + self.bwd_grad_act_net_time

    @property
    def net_time(self):
        return self.fwd_net_time + self.bwd_net_time + self.recompute_net_time

    @property
    def all_time(self):
        return self.fwd_time + self.fwd_net_time + self.bwd_time + self.bwd_net_time

    def get_all_costs(self):
        return [self.fwd_time, self.bwd_grad_act_time, self.bwd_grad_w_time]

    def __add__(self, other: "CostInfo") -> "CostInfo":
        return _add_dataclass(self, other)


@dataclass
class MemoryInfo:
    """Static (weights / main grads / optimizer state) bytes of a module, dense and MoE parts apart."""
    weight_numel: float = 0
    dense_weight_bytes: float = 0
    dense_grad_bytes: float = 0
    dense_state_bytes: float = 0
    moe_weight_numel: float = 0
    moe_weight_bytes: float = 0
    moe_grad_bytes: float = 0
    moe_state_bytes: float = 0
    # TE allocates one dummy wgrad buffer per distinct (rows, cols, elem_size): a *set*, not a sum
    te_dummy_wgrad_shapes: Set[Tuple[int, int, int]] = field(default_factory=set)

    @property
    def te_dummy_wgrad_bytes(self):
        return sum(r * c * e for r, c, e in self.te_dummy_wgrad_shapes)

    @property
    def all(self):
        return (self.dense_weight_bytes + self.dense_grad_bytes + self.dense_state_bytes + self.moe_weight_bytes
                + self.moe_grad_bytes + self.moe_state_bytes + self.te_dummy_wgrad_bytes)

    @property
    def all_state_bytes(self):
        return self.dense_state_bytes + self.moe_state_bytes

    @property
    def all_weight_bytes(self):
        return self.dense_weight_bytes + self.moe_weight_bytes

    @property
    def all_weight_numel(self):
        return self.weight_numel + self.moe_weight_numel

    @property
    def all_grad_bytes(self):
        return self.dense_grad_bytes + self.moe_grad_bytes

    def __add__(self, other: "MemoryInfo") -> "MemoryInfo":
        out = MemoryInfo()
        for f in fields(self):
            if f.name == "te_dummy_wgrad_shapes":
                out.te_dummy_wgrad_shapes = self.te_dummy_wgrad_shapes | other.te_dummy_wgrad_shapes
            else:
                setattr(out, f.name, getattr(self, f.name) + getattr(other, f.name))
        return out


@dataclass
class ActivationInfo:
    activation_mem_cache: float = 0
    fwd_peak_mem_no_cache: float = 0
    bwd_peak_mem_no_cache: float = 0
    cache_for_bwd_mem: float = 0     # scratch used by the lifetime walker

    @property
    def total_activation_mem_cache(self):
        return self.activation_mem_cache

    @property
    def fwd_peak_mem(self):
        return self.fwd_peak_mem_no_cache

    @property
    def bwd_peak_mem(self):
        return self.bwd_peak_mem_no_cache


_STAGES = ("forward", "backward", "recompute_forward", "recompute_backward")


@dataclass
class PeakPoint:
    """Peak live activation memory of one micro-batch through a stage chunk, per phase."""
    fwd_peak_path: Optional[str] = None
    fwd_peak_mem: float = 0.0
    bwd_peak_path: Optional[str] = None
    bwd_peak_mem: float = 0.0
    recomp_fwd_peak_path: Optional[str] = None
    recomp_fwd_peak_mem: float = 0.0
    recomp_bwd_peak_path: Optional[str] = None
    recomp_bwd_peak_mem: float = 0.0
    forward_activation_mem_cache: float = 0.0
    cur_stage: str = "forward"

    _FIELD = {"forward": "fwd", "backward": "bwd", "recompute_forward": "recomp_fwd", "recompute_backward": "recomp_bwd"}

    def set_stage(self, stage: str) -> None:
        assert stage in _STAGES
        self.cur_stage = stage

    def set_peak(self, path, mem, stage: str) -> None:
        self.set_stage(stage)
        prefix = self._FIELD[stage]
        setattr(self, prefix + "_peak_path", path)
        setattr(self, prefix + "_peak_mem", mem)

    def update_peak(self, path, mem, stage: str) -> None:
        if mem >= self.peak_mem:          # ">=": later equal peaks win (affects only the reported path)
            self.set_peak(path, mem, stage)

    def set_forward_mem_cache(self, mem_cache) -> None:
        self.forward_activation_mem_cache = mem_cache

    @property
    def activation_mem_cache(self):
        return self.forward_activation_mem_cache

    @property
    def peak_mem(self):
        return max(self.fwd_peak_mem, self.bwd_peak_mem, self.recomp_fwd_peak_mem, self.recomp_bwd_peak_mem)

    def _argmax_prefix(self) -> str:
        pk = self.peak_mem
        for stage in _STAGES[:3]:
            if pk == getattr(self, self._FIELD[stage] + "_peak_mem"):
                return stage
        return "recompute_backward"

    @property
    def peak_stage(self) -> str:
        return self._argmax_prefix()

    @property
    def peak_path(self):
        return getattr(self, self._FIELD[self._argmax_prefix()] + "_peak_path")

    def to_dict(self):
        d = asdict(self)
        d["activation_mem_cache"] = self.activation_mem_cache
        d["peak_stage"] = self.peak_stage
        d["peak_path"] = self.peak_path
        d["peak_mem"] = self.peak_mem
        del d["cur_stage"]
        del d["forward_activation_mem_cache"]
        return d

    def __repr__(self):
        return f"PeakPoint(path={self.peak_path}, peak_mem={self.peak_mem / 1024 ** 3:.4f} GB, peak_stage={self.peak_stage})"


class Result:
    """Thin wrapper around a result dict (reference model_struct.py:415)."""

    def __init__(self, result: dict) -> None:
        self.data = result

    def get(self, key: str):
        return self.data.get(key, None)

    def __getitem__(self, key):
        return self.data[key]

    def __contains__(self, key):
        return key in self.data
