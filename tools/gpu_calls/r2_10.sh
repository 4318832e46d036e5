set -uo pipefail
OUT=gpurun_out
echo "== attention tests (split 2)"
timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu -x 2>&1 | tail -4
echo "== fwd bench split2"; timeout 120 python tools/fmha_bench.py 2>&1 | tail -6 | cut -c1-230
echo "== fwd bench split1"; SMX_FMHA_SPLIT=1 timeout 120 python tools/fmha_bench.py 2>&1 | tail -6 | cut -c1-230
