#!/bin/bash
# the one-GPU records of a round: default bench line, reference arm, 2^31 and 2^32 points, smoke, exchange kernel profile
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 600 gpurun_out/r2_bench_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; cat gpurun_out/r2_bench_reference.json | cut -c1-400
python bench.py --log2n 31 --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_1gpu_2p31.json 2>> gpurun_out/r2_bench_1gpu.err
python bench.py --strong --steps 3 --no-e2e --no-cpu-baseline --no-alt > gpurun_out/r2_bench_1gpu_strong_2p32.json 2>> gpurun_out/r2_bench_1gpu.err
python bench.py --key-mask 0xFFFFFF --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_1gpu_mask24.json 2>> gpurun_out/r2_bench_1gpu.err
python bench.py --radix 11 --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_1gpu_radix11.json 2>> gpurun_out/r2_bench_1gpu.err
python bench.py --radix 8 --steps 3 --no-e2e --no-cpu-baseline --no-alt > gpurun_out/r2_bench_1gpu_radix8.json 2>> gpurun_out/r2_bench_1gpu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"exchange_vr_kernel" -c 2 -o gpurun_out/prof_r2_exchange_local_2p28 python tools/prof_sort.py --log2n 28 --iters 1 --no-skip --two-level > gpurun_out/r2_ncu_exchange.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_1gpu*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 3), "alt", d.get("alt_path", {}).get("ms_per_step"), "skip", d.get("skip_variant"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r2_bench_1gpu.err
