#!/bin/bash
# Build the library here (nvcc cross-compiles sm_100a without a GPU), then run a command on a B200 box.
#   tools/gpu.sh [--timeout S] [--gpus N] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py >/dev/null
exec /usr/local/graft/bin/gpurun "$@"
