#!/bin/bash
# One-GPU profile captures for profiles/ (run under gpurun; each capture only after the plain run exited 0).
cd ${GRAFT_REPO_ROOT:-.}
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
python tools/prof_sort.py --log2n 30 --iters 1 --no-skip > gpurun_out/r2_plain30.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"partition_kernel|subdigit_hist" -c 3 \
    -o gpurun_out/prof_r2_two_step_2p30 python tools/prof_sort.py --log2n 30 --iters 1 --no-skip > gpurun_out/r2_ncu_two_step.log 2>&1
python tools/prof_sort.py --log2n 28 --iters 1 --no-skip --one-pass > gpurun_out/r2_plain28_onepass.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"onepass_kernel|digit_hist" -c 2 \
    -o gpurun_out/prof_r2_one_pass_2p28 python tools/prof_sort.py --log2n 28 --iters 1 --no-skip --one-pass > gpurun_out/r2_ncu_one_pass.log 2>&1
# the virtual-rank pass shape on one GPU (count per part, scans, part sort, exchange kernel with local stores only)
python tools/prof_sort.py --log2n 28 --iters 1 --no-skip --two-level > gpurun_out/r2_plain28_two_level.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"exchange_vr_kernel|digit_hist_kernel" -c 3 \
    -o gpurun_out/prof_r2_two_level_2p28 python tools/prof_sort.py --log2n 28 --iters 1 --no-skip --two-level > gpurun_out/r2_ncu_two_level.log 2>&1
ls -la gpurun_out | tail -20
