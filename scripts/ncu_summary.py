"""Summarise .ncu-rep files (read here, no GPU needed) into a small text file for profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        print("#", path)
        for r in rows[2:]:
            print("kernel:", r[hdr.index("Kernel Name")], " id:", r[0])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    print("   %-68s %s %s" % (k, r[i], units[i]))
            st = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                    try:
                        st.append((float(r[i]), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
                    except ValueError:
                        pass
            tot = sum(a for a, _ in st) or 1.0
            st.sort(reverse=True)
            print("   stall samples: " + ", ".join("%s %.0f%%" % (b, 100 * a / tot) for a, b in st[:6]))
        print()


if __name__ == "__main__":
    main()
