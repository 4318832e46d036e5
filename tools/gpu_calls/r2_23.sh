set -uo pipefail
mkdir -p gpurun_out
timeout 200 python tools/profile_step.py --zoo-case deepseekv2_l2_ep1_pp1_dp1_mbc2 2>&1 | tail -34
