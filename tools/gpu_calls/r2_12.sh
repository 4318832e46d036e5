# 8 GPUs: bench both arms, predicted-vs-measured rows of the published table, correctness at 8, tp8 step profile
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
N=8
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port "$2" "${@:3}" 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm_mem\|Recompute is\|warnings.warn\|Gracefully"; }
echo "== bench N=8 own"; run 300 29725 bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/bench_r2_n8.json 2> $OUT/bench_r2_n8.err; tail -c 600 $OUT/bench_r2_n8.err; head -c 1200 $OUT/bench_r2_n8.json; echo
echo "== bench N=8 reference"; run 300 29726 bench.py --impl reference --gpus 8 --steps 3 --warmup 3 > $OUT/bench_r2_ref_n8.json 2> $OUT/bench_r2_ref_n8.err; tail -c 600 $OUT/bench_r2_ref_n8.err; head -c 1000 $OUT/bench_r2_ref_n8.json; echo
echo "== perf vs real (8-GPU rows)"; run 400 29727 tools/b200/perf_vs_real.py --out $OUT/perf_vs_real_8gpu --cases llama3_70b_l12_tp2_pp1_dp4_mbc8_selective llama3_70b_l12_tp1_pp2_dp4_mbc4 llama3_8b_tp1_pp2_dp4_mbc8 llama3_70b_l12_tp1_cp8_seq32768_mbc4 | tail -6
echo "== multi check"; run 200 29721 tools/multi_check.py | tail -9 | cut -c1-300
echo "== deepseek check"; run 120 29724 tools/deepseek_check.py | tail -1
echo "== tp8 step profile"; run 120 29728 tools/profile_step.py | tail -32
