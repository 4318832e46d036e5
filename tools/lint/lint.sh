#!/usr/bin/env bash
# Static checks. pylint is optional in the image; pyflakes-level checking via python -m compileall always runs.
set -euo pipefail
cd "$(dirname "$0")/../.."
python -m compileall -q simumax_b200 tools examples app bench.py __graft_entry__.py
if python -c "import pylint" 2>/dev/null; then
  python -m pylint --rcfile tools/lint/pylintrc simumax_b200
else
  echo "pylint not installed; compileall passed"
fi
