#!/bin/bash
# usage (under gpurun --gpus N): tools/run_final_mgpu.sh N "suite string"
# 1. bit-exact multi-GPU parity (tests/test_multigpu.py for world N), 2. the bench suite in one torchrun launch.
N=$1; SUITE=$2
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_mgpu${N}_gpus.txt
timeout 900 python -m pytest tests/test_multigpu.py -q -k "parity[$N]" > gpurun_out/r2_multigpu_parity_${N}gpu.log 2>&1
tail -3 gpurun_out/r2_multigpu_parity_${N}gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $N --steps 3 --warmup 3 --suite "$SUITE" > gpurun_out/r2_suite_${N}gpu.jsonl 2> gpurun_out/r2_suite_${N}gpu.err
python - <<PY
import json
for line in open("gpurun_out/r2_suite_${N}gpu.jsonl"):
    line = line.strip()
    if not line.startswith("{"): continue
    d = json.loads(line)
    if "error" in d: print(d); continue
    print(d["suite"], "value %.0f M/s" % d["value"], "ms %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], d["roofline"]["bound"], d.get("skip_variant"))
PY
tail -3 gpurun_out/r2_suite_${N}gpu.err
