"""The reference's published B200 table (docs/b200/b200_release_v1.2_summary.json: 43 rows, Llama-3-70B / 405B, DeepSeek-V2 / V3, TP / PP /
EP, CP at 32 K and 128 K) replayed on its shipped system config: every `perf_ms` and `perf_alloc_gib` must come out as recorded.
(tools/replay_published_table.py also runs the reference code next to it; skipped when no reference checkout is mounted.)"""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
SUMMARY = os.path.join(REF, "docs", "b200", "b200_release_v1.2_summary.json")
pytestmark = pytest.mark.skipif(not os.path.exists(SUMMARY), reason="reference checkout (published table + configs) not available")
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.timeout(900)
def test_every_published_row_is_reproduced():
    import replay_published_table as rp
    rows = json.load(open(SUMMARY))["rows"]
    assert len(rows) == 43
    bad = []
    for row in rows:
        ms, gib, misses = rp.run("simumax_b200", row)
        if abs(ms - row["perf_ms"]) > max(0.011, 2e-4 * row["perf_ms"]) or abs(gib - row["perf_alloc_gib"]) > 0.011 or misses:
            bad.append((row["case"], ms, row["perf_ms"], gib, row["perf_alloc_gib"], misses))
    assert not bad, bad


@pytest.mark.timeout(300)
def test_sync_vpp_preview_row_perf_and_simulator():
    """docs/release_v1.2.md:33-35: llama3_70b_l8_tp1_pp4_vp2_dp2_mbc8_sync -- perf 661.21 ms, simulator trace end 663.29 ms."""
    import replay_published_table as rp
    assert rp.vpp_preview() == (661.21, 663.29)
