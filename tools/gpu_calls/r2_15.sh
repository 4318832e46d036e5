# 1 GPU: re-measure the attention keys (second-generation forward, group-split backward) into the v6 sweep (supplement mode)
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
cp profiles/calib_ops_v6_without_attention.json $OUT/calib_ops_v7.json
timeout 600 python -m simumax_b200.calib.calibrate ops --default-sweep --sustained-ms 100 --cache-mode supplement --out $OUT/calib_ops_v7.json > $OUT/calib_v7.log 2>&1; echo rc=$?; grep -c "sdp_" $OUT/calib_v7.log; tail -4 $OUT/calib_v7.log
echo "== deepseek (shared MoE buffers)"; timeout 200 python tools/deepseek_check.py 2>&1 | tail -1
timeout 300 python tools/b200/perf_vs_real.py --out $OUT/perf_vs_real_ds1 --cases deepseekv2_l2_ep1_pp1_dp1_mbc2 llama3_8b_l16_selective_mbc4 2>&1 | grep -v Warn | tail -3
