"""tools/ncu_traffic.py REP GRID [REP GRID ...] -- DRAM bytes per launch of the graded kernel from `ncu --set full` captures
(dram__bytes_read.sum + dram__bytes_write.sum of the captured launch), written to profiles/step3d_t_traffic.json with the commit
the library was built from; bench.py reports it as roofline.traffic.   e.g.
    python tools/ncu_traffic.py gpurun_out/r2f_v8_b3.ncu-rep 2048x256x30"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(vals[ix[m]].replace(",", "")) * UNITS[units[ix[m]]]
    return tot, vals[ix["Kernel Name"]], float(vals[ix["gpu__time_duration.sum"]].replace(",", ""))


def main():
    path = os.path.join(ROOT, "profiles", "step3d_t_traffic.json")
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    for rep, grid in zip(sys.argv[1::2], sys.argv[2::2]):
        b, kern, dur = dram_bytes(rep)
        Lm, Mm, N = (int(x) for x in grid.split("x"))
        data[grid] = {"dram_bytes": b, "algorithmic_bytes": 96 * Lm * Mm * N, "ratio": b / (96.0 * Lm * Mm * N), "kernel": kern,
                      "source": "ncu --set full --clock-control none, one launch, %s, library built at or after commit %s" % (os.path.basename(rep), commit)}
        print(grid, data[grid])
    json.dump(data, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
