"""Attribute an ncu SASS source page (ncu -i rep --page source --csv) to CUDA source lines using
nvdisasm -g line markers of the same cubin (instructions are matched by order).
usage: ncu_lines.py <sass.csv> <cubin> <kernel-substring> [top]"""
import csv, re, subprocess, sys
sass_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []; cur = None; inside = False
for l in dis:
    if l.startswith("//---------------------"):
        inside = (kname in l) and ".text." in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 5]
print("sass rows", len(data), "disasm instrs", len(lines))
agg = {}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for n, r in enumerate(data):
    key = lines[n] if n < len(lines) else None
    a = agg.setdefault(key, {"inst": 0, "samples": 0, "st": {}})
    a["inst"] += int(r[ix["Instructions Executed"]] or 0)
    a["samples"] += int(r[ix["# Samples"]] or 0)
    for h in stalls:
        v = r[ix[h]]
        if v and v != "0": a["st"][h] = a["st"].get(h, 0) + int(v)
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samples"] for a in agg.values())
print("total inst %d samples %d" % (ti, ts))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(a["st"].items(), key=lambda kv: -kv[1])[:2]
    print("%-28s inst %5.1f%%  samples %5.1f%%  %s" % (key, 100.0 * a["inst"] / ti, 100.0 * a["samples"] / ts, st))
