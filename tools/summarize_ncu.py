"""Key metrics of an `ncu --set full` report, one block per captured launch (run where ncu is installed):
    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_ncu_full.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__cluster_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_sectors.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} launch(es); ncu --set full --clock-control none")
    for n, r in enumerate(rows[2:]):
        print(f"\n## launch {n}")
        for k in KEYS:
            if k in ci and r[ci[k]] != "":
                print(f"{k:75s} {r[ci[k]]} {units[ci[k]]}")
        for h, i in ci.items():
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and r[i] not in ("", "0"):
                print(f"{h:75s} {r[i]}")


if __name__ == "__main__":
    main()
