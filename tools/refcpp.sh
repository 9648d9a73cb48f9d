#!/bin/bash
# Reading aid: cpp-expand a reference source for one option set so only the
# ACTIVE Fortran is visible (SURVEY.md §0 recipe).  Output goes to /tmp, never
# into this repo.  usage: tools/refcpp.sh bench|upw [mpi] <path relative to /root/reference>
cfg=$1; shift
extra=""
if [ "$1" = mpi ]; then extra="-DMPI"; shift; fi
src=$1
cd /root/reference || exit 1
if [ "$cfg" = bench ]; then defs=(-DBENCHMARK '-DROMS_HEADER="benchmark.h"'); else defs=(-DUPWELLING -DROMS_HEADER="\"/tmp/refcpp/inc/upwelling_nodiag.h\"" ); fi
cpp -traditional -w -P -IROMS/Include -IROMS/Nonlinear -IROMS/Utility -IROMS/Functionals -IROMS/Modules "${defs[@]}" $extra -DNestedGrids=1 "$src" | grep -v '^\s*$' | grep -v '^!'
