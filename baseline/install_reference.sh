#!/usr/bin/env bash
# One-time offline install of the UNMODIFIED reference into baseline/_ref (git-ignored, travels with gpurun).
# The reference's setup.py lists only the top-level package (packages=['simumax']), so the stock pip install ships
# simumax/__init__.py, utils.py, version.py and nothing else; the sub-packages and the JSON configs its examples read
# are therefore copied verbatim next to it (no file is edited).
set -euo pipefail
cd "$(dirname "$0")/.."
rm -rf /tmp/ref_copy && cp -r /root/reference /tmp/ref_copy
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /tmp/ref_copy
for sub in core pp_simu testing tuning; do
  rm -rf "baseline/_ref/simumax/$sub" && cp -r "/root/reference/simumax/$sub" "baseline/_ref/simumax/$sub"
done
rm -rf baseline/_ref/configs && cp -r /root/reference/configs baseline/_ref/configs
find baseline/_ref -name __pycache__ -prune -exec rm -rf {} +
python - <<'PY'
import sys; sys.path.insert(0, "baseline/_ref")
from simumax.core.perf_llm import PerfLLM
print("reference importable:", PerfLLM.__module__)
PY
