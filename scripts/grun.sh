#!/bin/bash
# build the library here (nvcc cross-compiles), then run the given command on a B200 box:  scripts/grun.sh [--gpus N] [--timeout S] -- '<cmd>'
set -e
cd "$(dirname "$0")/.."
python -m rosnet_b200.csrc.build > /dev/null
exec /usr/local/graft/bin/gpurun "$@"
