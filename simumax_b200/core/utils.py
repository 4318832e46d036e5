"""Formatting, rank-layout and small shared helpers.

* human-readable conversion keeps the reference's output conventions ("1.0066 s", "50.8854 GB",
  "843.3426 T"; reference core/utils.py:47-170) because downstream tools parse those strings.
* rank layout: dense order tp -> cp -> dp -> pp (tp fastest), MoE order ep -> edp over the same
  global rank (reference core/utils.py:215-249).  Group-id strings are rendezvous keys in the simulator.
"""
from __future__ import annotations

import json
from typing import Dict

BYTE_UNITS = ["B", "KB", "MB", "GB", "TB"]
NUM_UNITS = ["", "K", "M", "B", "T"]
TIME_UNITS = ["ms", "s"]


def _scale(value, base, units, target_unit=None):
    v = float(value)
    if target_unit is not None:
        return v / (base ** units.index(target_unit)), target_unit
    i = 0
    while v >= base and i < len(units) - 1:
        v /= base
        i += 1
    return v, units[i]


def human_readable_bytes(value, target_unit=None) -> str:
    v, u = _scale(value, 1024, BYTE_UNITS, target_unit)
    return f"{v:.4f} {u}"


def human_readable_nums(value, target_unit=None) -> str:
    v, u = _scale(value, 1000, NUM_UNITS, target_unit)
    return f"{v:.4f} {u}"


def human_readable_times(value, target_unit=None) -> str:
    v, u = _scale(value, 1000, TIME_UNITS, target_unit)
    return f"{v:.4f} {u}"


class HumanReadableSize:
    """Parse / convert "<value> <unit>" strings (kept for scripts that consume formatted results)."""

    BYTE_UNITS = BYTE_UNITS
    NUM_UNITS = NUM_UNITS
    TIME_UNITS = TIME_UNITS

    def __init__(self, value, base=1024, units=None, source_unit=None, target_unit=None):
        self.units = units or ["B", "KB", "MB", "GB", "TB", "PB"]
        self.base = base
        src = source_unit or self.units[0]
        raw = float(value) * (base ** self.units.index(src))
        self.converted_value, self.unit = _scale(raw, base, self.units, target_unit)

    @staticmethod
    def from_string(size_str, units, base, target_unit=None):
        value, unit = size_str.split(" ")
        if unit not in units:
            raise ValueError(f"Unknown unit: '{unit}'")
        return HumanReadableSize(float(value), base=base, units=units, source_unit=unit, target_unit=target_unit)

    def get_value(self):
        return self.converted_value

    def get_unit(self):
        return self.unit

    def __str__(self):
        return f"{self.converted_value:.4f} {self.unit}"


def parse_time_ms(text) -> float:
    """'1.0066 s' / '663.7697 ms' -> milliseconds (numbers pass through)."""
    if isinstance(text, (int, float)):
        return float(text)
    return HumanReadableSize.from_string(text, TIME_UNITS, 1000, target_unit="ms").get_value()


def parse_bytes(text) -> float:
    if isinstance(text, (int, float)):
        return float(text)
    return HumanReadableSize.from_string(text, BYTE_UNITS, 1024, target_unit="B").get_value()


def convert_final_result_to_human_format(result: dict) -> None:
    """In place: keys containing time -> ms/s, mem|bytes -> B..TB, flops -> K/M/B/T (key-name heuristic)."""
    if result is None:
        return
    for k, v in result.items():
        if isinstance(v, dict):
            convert_final_result_to_human_format(v)
            continue
        if not isinstance(v, (int, float)):
            continue
        if "time" in k:
            result[k] = human_readable_times(v)
        elif "mem" in k or "bytes" in k:
            result[k] = human_readable_bytes(v)
        elif "flops" in k:
            result[k] = human_readable_nums(v)


def to_json_string(obj: dict) -> str:
    return json.dumps(obj, indent=2, sort_keys=False, ensure_ascii=False)


def path_convert_to_str(path: list) -> str:
    return " -> ".join(path) if path else ""


def get_point_name(parent, current, sep=" -> ") -> str:
    if parent and current:
        return parent + sep + current
    return parent if parent else current


# ------------------------------------------------------------------------------------------------
# rank layout
# ------------------------------------------------------------------------------------------------
def get_pp_stage_representative_rank(pp_rank: int, strategy) -> int:
    """tp=0, cp=0, dp=0 rank of a pipeline stage."""
    return pp_rank * strategy.tp_size * strategy.cp_size * strategy.dp_size


def get_pp_p2p_comm_size(strategy, hidden_size, dtype_size):
    size = strategy.micro_batch_size * strategy.seq_len * hidden_size * dtype_size / strategy.cp_size
    if strategy.enable_sequence_parallel:
        size = size / strategy.tp_size
    return size


def get_rank_group(global_rank: int, strategy) -> Dict[str, object]:
    tp, cp, dp = strategy.tp_size, strategy.cp_size, strategy.dp_size
    tp_rank = global_rank % tp
    cp_rank = (global_rank // tp) % cp
    dp_rank = (global_rank // (tp * cp)) % dp
    dp_cp_rank = (global_rank // tp) % (cp * dp)
    pp_rank = global_rank // (tp * cp * dp)
    ep_rank = global_rank % strategy.ep_size
    edp_rank = global_rank // strategy.ep_size % strategy.edp_size
    return {
        "tp_group_id": f"pp:{pp_rank}-cp:{cp_rank}-dp:{dp_rank}", "tp_rank": tp_rank,
        "cp_group_id": f"tp:{tp_rank}-pp:{pp_rank}-dp:{dp_rank}", "cp_rank": cp_rank,
        "pp_group_id": f"tp:{tp_rank}-cp:{cp_rank}-dp:{dp_rank}", "pp_rank": pp_rank,
        "dp_group_id": f"tp:{tp_rank}-pp:{pp_rank}", "dp_rank": dp_rank,
        "dp_cp_group_id": f"tp:{tp_rank}-pp:{pp_rank}", "dp_cp_rank": dp_cp_rank,
        "ep_group_id": f"tp:{tp_rank}-pp:{pp_rank}-edp:{edp_rank}", "ep_rank": ep_rank,
        "edp_group_id": f"tp:{tp_rank}-pp:{pp_rank}-ep:{ep_rank}", "edp_rank": edp_rank,
    }


def format_scope_microbatch_tag(args, include_chunk: bool = False) -> str:
    tag = f"microbatch{args.microbatch}"
    chunk_idx = getattr(args, "chunk_idx", None)
    if include_chunk and chunk_idx is not None:
        tag += f"-chunk{chunk_idx}"
    return tag


def format_model_info_microbatch_tag(args) -> str:
    tag = f"microbatch:{args.microbatch}"
    chunk_idx = getattr(args, "chunk_idx", None)
    if chunk_idx is not None:
        tag += f"-chunk:{chunk_idx}"
    return tag
