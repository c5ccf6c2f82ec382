"""Selected metrics of the kernels in an .ncu-rep (ncu --set full) as a small "kernel","metric","unit","value" CSV for
profiles/.  Usage (here, no GPU needed): python tools/ncu_rep_to_csv.py gpurun_out/x.ncu-rep profiles/x.csv"""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.per_cycle_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow(["kernel", "metric", "unit", "value"])
        for vals in rows[2:]:
            kname = vals[name_col].split("(")[0].split("::")[-1]
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP or h.startswith("smsp__average_warps_issue_stalled"):
                    w.writerow([kname, h, u, v])
    print("wrote", out)


if __name__ == "__main__":
    main()
