#!/bin/bash
# build in-tree, then run a command on the GPU box:  scripts/gpu.sh <timeout> '<command>'
set -e
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; exit 1; }
/usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
