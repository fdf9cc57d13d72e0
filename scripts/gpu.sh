#!/usr/bin/env bash
# local helper: (re)build the CUDA library, check that it exports every declared symbol, then run scripts/_gpu_call.sh on the GPU box
#   scripts/gpu.sh [gpurun flags, e.g. --gpus 2]   -> log in /tmp/gpu_last.log
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" | tail -2
gpurun "$@" --timeout ${GPU_TIMEOUT:-1800} -- 'bash scripts/_gpu_call.sh' > /tmp/gpu_last.log 2>&1
tail -3 /tmp/gpu_last.log
