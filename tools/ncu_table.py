"""tools/ncu_table.py REP.ncu-rep [...] -- one line per report: duration, DRAM bytes, throughputs, registers, occupancy, cache hit
rates, top stall reasons (warps stalled per issued instruction).  Reads the reports with `ncu -i ... --page raw --csv` (run where ncu is installed; no GPU needed)."""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
SCALE = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1e3, "msecond": 1e3, "us": 1.0, "usecond": 1.0, "ns": 1e-3, "nsecond": 1e-3, "second": 1e6}


def row(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    h, u, v = r[0], r[1], r[2]
    d = {h[i]: (v[i], u[i]) for i in range(len(h))}

    def g(n):
        if n not in d or d[n][0] in ("", "n/a"):
            return float("nan")
        return float(d[n][0].replace(",", "")) * SCALE.get(d[n][1], 1.0)
    stalls = sorted(((float(d[k][0].replace(",", "")), k.split("issue_stalled_")[1].split("_per_issue")[0]) for k in d
                     if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k
                     and d[k][0] not in ("", "n/a")), reverse=True)[:3]
    name = d.get("Kernel Name", ("?", ""))[0].split("(")[0]
    print("%-24s %8.1f us  dram rd %7.1f wr %7.1f MB  dram %5.1f%%  sm %5.1f%%  fp64 %5.1f%%  issue %5.1f%%  regs %3d  occ %5.1f%%  L1 %5.1f  L2 %5.1f  stalls: %s"
          % (name, g(WANT[0]), g(WANT[1]), g(WANT[2]), g(WANT[8]), g(WANT[7]), g(WANT[9]), g(WANT[10]), int(g(WANT[3])), g(WANT[4]), g(WANT[5]), g(WANT[6]),
             ", ".join("%s %.1f" % (n, x) for x, n in stalls)))


for p in sys.argv[1:]:
    row(p)
