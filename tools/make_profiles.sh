#!/bin/bash
# tools/make_profiles.sh TAG -- profiles/r02_* from the files tools/r2_evidence.sh TAG left in gpurun_out/ (run here; ncu reads the reports without a GPU)
cd "$(dirname "$0")/.."; G=gpurun_out/$1; H=$(git rev-parse --short HEAD)
cp ${G}_phases_b1.log profiles/r02_phase_times_bench1.txt; cp ${G}_phases_b3.log profiles/r02_phase_times_bench3.txt
python - "$G" <<'PY'
import csv, collections, sys
G=sys.argv[1]
rows=list(csv.DictReader(l for l in open(G+'_launches.csv') if l.startswith('"')))
agg=collections.OrderedDict()
for r in rows:
    k=r["Kernel Name"].split("(")[0].replace("void ","")
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r["Metric Value"].replace(",",""))/1e3
tot=sum(v[1] for v in agg.values())
out=["ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 python bench.py --steps 4 --warmup 3 --no-cpu --no-roofline --no-check",
     "400 consecutive launches inside the timed region (BENCHMARK1, 1 GPU; cold-cache, serialised replays: SHARES, not absolute times)",
     "%-44s %6s %10s %7s"%("kernel","count","sum us","share")]
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    out.append("%-44s %6d %10.1f %6.1f%%"%(k,n,t,100*t/tot))
out.append("%-44s %6d %10.1f"%("total",sum(v[0] for v in agg.values()),tot))
open('profiles/r02_bench_launches.txt','w').write("\n".join(out)+"\n")
PY
{ echo "ncu --set full --clock-control none, ONE warm launch per kernel on 2048x256x30 (15.7 M cells, one B200), read with tools/ncu_table.py"
  echo "(stalls = warps stalled for that reason per issued instruction).  Round-2 kernels (tools/r2_evidence.sh, commit $H):"
  python tools/ncu_table.py ${G}_t3dmix2_geo_roll_kernel.ncu-rep ${G}_pre_step3d_t_roll_kernel.ncu-rep ${G}_uv3dmix2_roll_kernel.ncu-rep ${G}_rhs3d_roll_kernel.ncu-rep ${G}_pre_step3d_uv_march_kernel.ncu-rep ${G}_step2d_kernel.ncu-rep ${G}_v8_b3.ncu-rep
  echo "step2d on BENCHMARK1 (512x64):"; python tools/ncu_table.py ${G}_step2d_b1.ncu-rep
  echo; echo "Per-level (round-1 form) kernels and the kernels not yet restructured (tools/ncu_kernels.sh, commit 9e8b9c4; this capture is what motivated the marching kernels):"
  python tools/ncu_table.py gpurun_out/r2s_*.ncu-rep
  echo; echo "pre_step3d_t_kernel (per-level form): L1 -> L2 read sectors 210 M (6.7 GB), DRAM read 148 M sectors (4.75 GB) for 1.5 GB of operands; long-scoreboard 10.5 stalled warps per issue, 7.3 warps per scheduler."
} > profiles/r02_ncu_kernels_2048x256x30.txt 2>&1
python tools/ncu_summary.py ${G}_v8_b3.ncu-rep > profiles/r02_step3d_t_v8_ncu_full_2048x256x30.txt 2>&1
python tools/ncu_traffic.py ${G}_v8_b3.ncu-rep 2048x256x30 > /dev/null
{ echo "ROMS_B200_S2_PERSIST=1 ROMS_B200_S2_PROF=1 python tools/time_phases.py   (BENCHMARK1 512x64, 289 blocks of 256 threads on 148 SMs)"
  echo "clock64 stamps of the middle block, cycles per sub-step at 1965 MHz:"; grep -h "persistent\|step2d_loop" ${G}_persist.log
  echo; echo "the same loop as a CUDA graph of per-sub-step launches chained by programmatic dependent launch (production):"; grep -h step2d_loop ${G}_phases_b1.log
  echo; echo "first cut of the persistent kernel (body of the per-sub-step kernel unchanged, no early loads, metrics from global memory):"
  grep -h step2d gpurun_out/r2p_phases_b1.log gpurun_out/r2p_phases_b1_graph.log; echo "(first line persistent, second line graph, same build)"; } > profiles/r02_step2d_persistent.txt
{ echo "bench.py lines of round 2 (commit $H for N=1; multi-GPU lines N=2, 4, 8: tools/mgpu_bench.sh / tools/mgpu_ab.sh, final build)"; echo "--- N=1"; tail -1 ${G}_bench.log; echo "--- N=1 --impl reference"; tail -1 ${G}_bench_ref.log
  for f in gpurun_out/fin2b_bench.log gpurun_out/r2ai_X=1.log gpurun_out/fin8b_bench.log; do [ -f $f ] && { echo "--- $f"; grep '"metric"' $f | tail -1; }; done; } > profiles/r02_bench_lines.txt
git rm -q --cached profiles/r02_multi_gpu_bench_lines.txt 2>/dev/null; rm -f profiles/r02_multi_gpu_bench_lines.txt
ls -la profiles | grep r02
