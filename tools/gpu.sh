#!/bin/bash
# dev helper: rebuild libscz.so + the oracle, then run a command on the GPU box
set -e
cd "$(dirname "$0")/.."
make -s -j8 -C scalable-collaborative-zksnark_b200 2>&1 | grep -v "^$" | tail -5
make -s -C oracle
T=${GPU_TIMEOUT:-900}
exec /usr/local/graft/bin/gpurun --timeout $T ${GPU_ARGS:-} -- "$@"
