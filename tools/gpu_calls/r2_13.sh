set -uo pipefail
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$2" "${@:3}" 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm_mem\|Recompute is\|warnings.warn\|Gracefully"; }
echo "== dp check"; run 200 29731 tools/dp_check.py | tail -2 | cut -c1-900
echo "== deepseek check"; run 200 29732 tools/deepseek_check.py | tail -1
echo "== pp check"; run 200 29733 tools/pp_check.py | tail -1
