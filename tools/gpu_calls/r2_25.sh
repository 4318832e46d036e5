set -uo pipefail
mkdir -p gpurun_out
timeout 120 python tools/profile_step.py --zoo-case llama3_70b_l4_mbc4 2>&1 | grep " ms \|sum of" | head -30
