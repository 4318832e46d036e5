# 4 GPUs: TP x DP / PP x DP / CP4 / EP4 correctness, bench N=4 both arms, predicted vs measured rows
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
N=4
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port "$2" "${@:3}" 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm_mem\|Recompute is\|warnings.warn\|Gracefully"; }
echo "== multi check"; run 300 29711 tools/multi_check.py | tail -9 | cut -c1-900
echo "== cp check"; run 200 29712 tools/cp_check.py | tail -2 | cut -c1-1800
echo "== pp check"; run 200 29713 tools/pp_check.py | tail -2 | cut -c1-900
echo "== deepseek check"; run 200 29714 tools/deepseek_check.py | tail -2 | cut -c1-400
echo "== bench N=4 own"; run 400 29715 bench.py --gpus 4 --steps 3 --warmup 3 > $OUT/bench_r2_n4.json 2> $OUT/bench_r2_n4.err; tail -c 800 $OUT/bench_r2_n4.err; head -c 1500 $OUT/bench_r2_n4.json; echo
echo "== bench N=4 reference"; run 400 29716 bench.py --impl reference --gpus 4 --steps 3 --warmup 3 > $OUT/bench_r2_ref_n4.json 2> $OUT/bench_r2_ref_n4.err; tail -c 800 $OUT/bench_r2_ref_n4.err; head -c 1200 $OUT/bench_r2_ref_n4.json; echo
echo "== perf vs real (4-GPU rows)"; run 600 29717 tools/b200/perf_vs_real.py --out $OUT/perf_vs_real_4gpu --cases llama3_70b_l12_tp2_pp1_dp2_mbc4 llama3_8b_tp1_pp2_dp2_mbc8 llama3_8b_tp1_pp4_dp1_mbc8 deepseekv2_l4_ep4_pp1_dp4_mbc4 llama3_8b_l16_tp1_cp4_seq32768_mbc2 | tail -8
