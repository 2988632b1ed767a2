"""Extract the counters the roofline discussion needs from an .ncu-rep (run where ncu is installed):

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_subpipe",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled",
    "sm__cycles_active.avg", "sm__cycles_active.max",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[name_col][:100]}  (id {r[0]})")
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(w) for w in WANT) and v not in ("", "n/a"):
                print(f"   {h} = {v} {u}")


if __name__ == "__main__":
    main(sys.argv[1])
