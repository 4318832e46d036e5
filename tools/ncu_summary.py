"""Condense an .ncu-rep into the handful of numbers worth committing (profiles/*.json).
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/x_ncu_summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__ops_path_tensor_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:120]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = float(r[i])
                except ValueError:
                    d[k] = r[i]
                d[k + "__unit"] = units[i]
        res.append(d)
    json.dump({"report": rep, "kernels": res}, open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"][:60], d.get("gpu__time_duration.sum"), "tensor%", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
