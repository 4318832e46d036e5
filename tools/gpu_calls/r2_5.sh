# 2 GPUs
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}" 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm_mem\|Recompute is\|warnings.warn"; }
echo "== multi check"; run 200 29701 tools/multi_check.py | tail -8 | cut -c1-1500
echo "== pp check"; run 200 29702 tools/pp_check.py | tail -3 | cut -c1-1500
echo "== deepseek check"; run 200 29703 tools/deepseek_check.py | tail -4 | cut -c1-2500
echo "== bench N=2 own"; run 400 29704 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_r2_n2.json 2> $OUT/bench_r2_n2.err; tail -c 1200 $OUT/bench_r2_n2.err; head -c 2500 $OUT/bench_r2_n2.json
echo "== bench N=2 reference"; run 400 29705 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > $OUT/bench_r2_ref_n2.json 2> $OUT/bench_r2_ref_n2.err; tail -c 1200 $OUT/bench_r2_ref_n2.err; head -c 2500 $OUT/bench_r2_ref_n2.json
