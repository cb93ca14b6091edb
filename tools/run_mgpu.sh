#!/bin/bash
# usage: tools/run_mgpu.sh NGPUS LOG2N "sets"   (under gpurun --gpus NGPUS)
cd ${GRAFT_REPO_ROOT:-.}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port 29541 tools/prof_mgpu.py --log2n $2 --iters 2 --sets "$3" 2>&1 | grep -E "^\[|rror|Traceback" | cut -c1-400
