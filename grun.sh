#!/bin/bash
# Rebuild every native artefact (the box has no nvcc-time for us: it runs the
# prebuilt in-tree .so files), then run a command on the B200 box.
# usage: ./grun.sh <timeout-seconds> '<command>'
set -e
cd "$(dirname "$0")"
python -c "import __graft_entry__ as g; g.build()" 1>&2
t=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
