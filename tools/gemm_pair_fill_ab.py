"""A/B of the pair-vs-single kernel choice on shapes with few 256 x 256 tiles (N = 256 ... 768 projections, 256 / 768-row wgrads):
old rule (pair whenever M, N >= 256) vs `_pair_fills_machine`, next to cuBLAS on the same operands, all at the sustained clock.
    python tools/gemm_pair_fill_ab.py [ops.json]   ->  gpurun_out/gemm_pair_fill_ab.json (+ the new numbers in calib-ops format)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from simumax_b200.calib import measure
from simumax_b200.calib.shapes import ShapeRequest, parse_key
from simumax_b200.ops.gemm import _pair_fills_machine, _pair_tiles_pay


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "profiles/calib_ops_v10_red_wgrad.json"
    table = json.load(open(src))["ops"]["matmul"]
    os.environ["SMX_GEMM_PAIR_FILL"] = "1"
    keys = []
    for k in table:
        f = parse_key(k)
        M = int(f["b"]) * int(f["m"]) if f["layout"] in ("TN", "NN") else int(f["m"])
        N = int(f["n"])
        if M >= 256 and N >= 256 and _pair_tiles_pay(M) and not _pair_fills_machine(1, M, N, 148):
            keys.append(k)
    measure.SUSTAINED_MS = 100.0
    rows, ops = [], {}
    for k in keys:
        req = ShapeRequest("matmul", k, table[k]["flops"], parse_key(k))
        measure.LIBRARY_COLUMN = False
        os.environ["SMX_GEMM_PAIR_FILL"] = "0"
        old = measure.measure_matmul(req)
        measure.LIBRARY_COLUMN = True
        os.environ["SMX_GEMM_PAIR_FILL"] = "1"
        new = measure.measure_matmul(req)
        rows.append({"key": k, "pair_tflops": old["tflops"], "single_tflops": new["tflops"], "cublas_tflops": new["cublas_tflops"],
                     "table_tflops": table[k]["tflops"]})
        new["efficiency"] = req.flops / (new["ms"] * 1e-3) / 2250e12
        ops[k] = new
        print(f"{k}: pair {old['tflops']:.0f} -> single {new['tflops']:.0f} TF (cuBLAS {new['cublas_tflops']:.0f}, table {table[k]['tflops']:.0f})", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"rows": rows, "ops": {"matmul": ops}}, open("gpurun_out/gemm_pair_fill_ab.json", "w"), indent=1)


if __name__ == "__main__":
    main()
