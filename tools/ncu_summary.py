"""Condense an .ncu-rep (ncu --set full) into the handful of metrics the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "sm clock"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (realtime)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "TMEM active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_scoreboard %"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall barrier %"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"kernel: {r[col['Kernel Name']][:110]}")
        for k, label in KEYS:
            if k in col:
                print(f"  {label:34s} {r[col[k]]:>16s} {units[col[k]]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
