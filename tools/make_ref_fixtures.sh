#!/bin/bash
# tools/make_ref_fixtures.sh ROMS_SRC [WORKDIR] -- reference-derived pins for tests/test_ref_fixtures.py.
#
# The reference (myroms/roms, Fortran) cannot be built in this repository's image (no Fortran compiler, no NetCDF-Fortran), so the
# oracle is pinned by the reference only through the equation-of-state check values.  On ANY box with gfortran
# and NetCDF-Fortran this script builds the unmodified reference for UPWELLING and BENCHMARK (serial, the compiler flags of
# Compilers/Linux-gfortran.mk:99-100 WITHOUT -ffast-math, so that the arithmetic is IEEE and the operation order the source's),
# runs 100 baroclinic steps of roms_upwelling.in and roms_benchmark1.in with NINFO=1, and turns the `diag` lines the reference prints
# every step (diag.F:472-500: step, avgke, avgpe, avgkp, volume in 1pe14.6, then the largest Courant number with its location and the
# maximum speed) into tests/golden/ref_upwelling.json / ref_benchmark1.json.  Commit those two files: the CPU test then checks the
# oracle against every printed digit of every step, and the parity label of the oracle changes from "unpinned" to "pinned".
# With ncdump available the 100-step history fields zeta,u,v,temp,salt (OUT_DOUBLE) are added as flat arrays.
set -euo pipefail
SRC=$(readlink -f "${1:?usage: make_ref_fixtures.sh ROMS_SRC [WORKDIR]}")
WORK=$(readlink -f "${2:-/tmp/roms_ref_fixtures}")
HERE=$(cd "$(dirname "$0")/.." && pwd)
command -v gfortran >/dev/null || { echo "gfortran not found"; exit 2; }
command -v nf-config >/dev/null || { echo "nf-config (NetCDF-Fortran) not found"; exit 2; }
mkdir -p "$WORK"
for APP in UPWELLING BENCHMARK; do
  app=$(echo $APP | tr A-Z a-z)
  D="$WORK/$app"; rm -rf "$D"; mkdir -p "$D"; cd "$D"
  cp "$SRC/ROMS/Bin/build_roms.sh" .
  # IEEE arithmetic: drop -ffast-math from the optimised flags (a private copy of the compiler file)
  mkdir -p Compilers; cp "$SRC"/Compilers/*.mk "$SRC"/Compilers/*.pl Compilers/ 2>/dev/null || true
  sed -i 's/^\( *FFLAGS += -ffast-math\)/#\1/' Compilers/Linux-gfortran.mk
  export MY_ROOT_DIR="$WORK" MY_PROJECT_DIR="$D" MY_ROMS_SRC="$SRC" COMPILERS="$D/Compilers"
  export ROMS_APPLICATION=$APP FORT=gfortran USE_NETCDF4=on
  unset USE_MPI USE_MPIF90 USE_OpenMP USE_DEBUG || true
  if [ "$APP" = BENCHMARK ]; then export MY_CPP_FLAGS="-DOUT_DOUBLE"; in=roms_benchmark1.in; else export MY_CPP_FLAGS="-DOUT_DOUBLE"; in=roms_upwelling.in; fi
  sed -i -e 's/^\( *export *USE_MPI=\)/#\1/' -e 's/^\( *export *USE_MPIF90=\)/#\1/' build_roms.sh
  bash build_roms.sh -j 8 > build.log 2>&1 || { tail -30 build.log; echo "build of $APP failed (see $D/build.log)"; exit 3; }
  cp "$SRC/ROMS/External/$in" . ; cp "$SRC/ROMS/External/varinfo.yaml" . 2>/dev/null || true
  sed -i -e 's/^\( *NTIMES *==\).*/\1 100/' -e 's/^\( *NINFO *==\).*/\1 1/' -e 's/^\( *NHIS *==\).*/\1 100/' -e 's/^\( *NDEFHIS *==\).*/\1 0/' \
         -e 's/^\( *NRST *==\).*/\1 1000/' -e 's/^\( *NAVG *==\).*/\1 1000/' -e 's/^\( *NDIA *==\).*/\1 1000/' \
         -e 's/^\( *NtileI *==\).*/\1 1/' -e 's/^\( *NtileJ *==\).*/\1 1/' -e "s#^\( *VARNAME *=\).*#\1 $SRC/ROMS/External/varinfo.yaml#" $in
  ./romsS < $in > roms.log 2>&1 || { tail -30 roms.log; echo "run of $APP failed"; exit 4; }
  name=$([ "$APP" = BENCHMARK ] && echo benchmark1 || echo upwelling)
  python3 - "$D/roms.log" "$HERE/tests/golden/ref_$name.json" "$D" <<'PY'
import glob, json, re, subprocess, sys
log, out, d = sys.argv[1:4]
steps = []
lines = open(log).read().splitlines()
num = r"[-+]?\d\.\d{6}E[-+]\d{2,3}"
for n, l in enumerate(lines):
    m = re.match(r"\s*(\d+)\s+\S+\s+\S+\s+(%s)\s*(%s)\s*(%s)\s*(%s)\s*$" % (num, num, num, num), l)
    if not m:
        continue
    rec = {"step": int(m.group(1)), "avgke": m.group(2), "avgpe": m.group(3), "avgkp": m.group(4), "volume": m.group(5)}
    c = re.match(r"\s*\((\d+),(\d+),(\d+)\)\s+(%s)\s+(%s)\s+(%s)\s+(%s)" % (num, num, num, num), lines[n + 1]) if n + 1 < len(lines) else None
    if c:
        rec.update({"Ci": int(c.group(1)), "Cj": int(c.group(2)), "Ck": int(c.group(3)), "Cu": c.group(4), "Cv": c.group(5), "Cw": c.group(6), "maxspeed": c.group(7)})
    steps.append(rec)
fix = {"source": "unmodified myroms/roms, gfortran -O3 without -ffast-math, serial, 100 steps, NINFO=1; numbers are the printed strings (1pe14.6)",
       "diag": steps}
his = sorted(glob.glob(d + "/*his*.nc"))
if his:
    try:
        for v in ("zeta", "u", "v", "temp", "salt"):
            txt = subprocess.run(["ncdump", "-p", "17", "-v", v, his[0]], capture_output=True, text=True, check=True).stdout
            body = txt[txt.index("data:"):]
            vals = re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?|_", body[body.index("=") + 1:body.index(";")])
            fix.setdefault("fields_last_record", {})[v] = [None if x == "_" else float(x) for x in vals]
    except Exception as e:      # noqa: BLE001
        fix["fields_error"] = repr(e)
json.dump(fix, open(out, "w"))
print("wrote", out, len(steps), "diag lines")
PY
done
echo "commit tests/golden/ref_upwelling.json and tests/golden/ref_benchmark1.json; python -m pytest tests/test_ref_fixtures.py"
