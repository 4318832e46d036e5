"""Golden-result comparison helpers: recursive dict diff with relative tolerance.

Used by the regression tests (tests/test_golden_*.py) to compare result dictionaries / JSON files against stored
goldens; numbers given as human strings ("1.0066 s", "50.8854 GB") are parsed before comparison.
(reference simumax/testing/base_test_tool.py)
"""
from __future__ import annotations

import json
import re
from typing import Any, List, Tuple

_NUM_UNIT = re.compile(r"^\s*(-?\d+(?:\.\d+)?(?:[eE][-+]?\d+)?)\s*([A-Za-z%]*)\s*$")
_SCALE = {"": 1.0, "ms": 1.0, "s": 1e3, "B": 1.0, "KB": 1024.0, "MB": 1024.0 ** 2, "GB": 1024.0 ** 3, "TB": 1024.0 ** 4,
          "K": 1e3, "M": 1e6, "T": 1e12, "%": 1.0}


def _to_number(x):
    if isinstance(x, bool):
        return None
    if isinstance(x, (int, float)):
        return float(x)
    if isinstance(x, str):
        m = _NUM_UNIT.match(x)
        if m and m.group(2) in _SCALE:
            return float(m.group(1)) * _SCALE[m.group(2)]
    return None


class RelDiffComparator:
    def __init__(self, rel_tol: float = 1e-4, abs_tol: float = 1e-9):
        self.rel_tol, self.abs_tol = rel_tol, abs_tol

    def __call__(self, a, b) -> bool:
        na, nb = _to_number(a), _to_number(b)
        if na is not None and nb is not None:
            return abs(na - nb) <= max(self.abs_tol, self.rel_tol * max(abs(na), abs(nb)))
        return a == b


class ResultCheck:
    """``ResultCheck(rel_tol).compare(result, golden)`` -> list of (path, got, expected) mismatches."""

    def __init__(self, rel_tol: float = 1e-4, ignore_keys=()):
        self.cmp = RelDiffComparator(rel_tol)
        self.ignore = set(ignore_keys)

    def compare(self, got: Any, expected: Any, path: str = "") -> List[Tuple[str, Any, Any]]:
        diffs: List[Tuple[str, Any, Any]] = []
        if isinstance(expected, dict) and isinstance(got, dict):
            for k in expected:
                if k in self.ignore:
                    continue
                if k not in got:
                    diffs.append((f"{path}/{k}", "<missing>", expected[k]))
                else:
                    diffs.extend(self.compare(got[k], expected[k], f"{path}/{k}"))
        elif isinstance(expected, (list, tuple)) and isinstance(got, (list, tuple)):
            if len(expected) != len(got):
                diffs.append((path + "/len", len(got), len(expected)))
            for i, (g, e) in enumerate(zip(got, expected)):
                diffs.extend(self.compare(g, e, f"{path}[{i}]"))
        elif not self.cmp(got, expected):
            diffs.append((path, got, expected))
        return diffs

    def check_files(self, got_path: str, golden_path: str):
        with open(got_path) as a, open(golden_path) as b:
            return self.compare(json.load(a), json.load(b))

    def assert_equal(self, got, expected):
        diffs = self.compare(got, expected)
        assert not diffs, "result mismatch:\n" + "\n".join(f"  {p}: got {g!r} expected {e!r}" for p, g, e in diffs[:20])
