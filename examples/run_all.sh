#!/usr/bin/env bash
# Runs every example in this directory (CPU only); stops at the first failure.
set -euo pipefail
cd "$(dirname "$0")"
for f in perf_*.py search_strategy_llama3_8b.py simulator_trace_snapshot.py show_simu_avaliable_modes.py; do
  echo "== $f"
  python "$f" > /dev/null
done
echo "all examples ran"
