set -uo pipefail
for d in 0 1 2 3; do echo "== diag $d"; SMX_FMHA_DIAG=$d timeout 100 python tools/fmha_bench.py 2>&1 | grep -v '"D": 192' | tail -4 | cut -c1-150; done
