#!/bin/bash
# tools/gpu.sh TIMEOUT 'command' -- rebuild everything in-tree (the .so travels with the snapshot), then run on a B200 box.
cd "$(dirname "$0")/.."
if python -c "import __graft_entry__ as g; g.build()" 2>&1 | grep -E " error |Error"; then echo "BUILD FAILED"; exit 1; fi
t=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
