set -uo pipefail
echo "== attention tests gen3"; SMX_FMHA_GEN=3 timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu -x 2>&1 | tail -4
echo "== fwd bench gen3"; SMX_FMHA_GEN=3 timeout 120 python tools/fmha_bench.py 2>&1 | grep -v '"D": 192' | tail -4 | cut -c1-230
echo "== fwd bench gen3 no cluster"; SMX_FMHA_GEN=3 SMX_FMHA_CLUSTER=1 timeout 120 python tools/fmha_bench.py 2>&1 | grep -v '"D": 192' | tail -4 | cut -c1-150
echo "== gen3 trace"; SMX_FMHA_GEN=3 timeout 100 python tools/fmha_fwd_trace.py 2>&1 | tail -10
