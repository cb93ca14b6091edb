#!/bin/bash
# One call on a 1-GPU box, most important step first (the GPU-minute budget may cut the tail):
#   1. tools/sweep_partition.py: A/B of partition_kernel's compile-time variants at n = 2^30, every sort verified;
#      the fastest one becomes the candidate default if it beats the library as built by > 1 %
#   2. the WHOLE `-m gpu` suite with the candidate applied through LSB_TEST_TUNE (tests/conftest.py)
#   3. smoke + the default bench line with the candidate applied through --tune
#   4. ncu launch list of the bench command, then one `--set full` capture of two scatter launches
# Outputs under gpurun_out/ (copied to profiles/r2_final_* afterwards).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/${TAG:-r2_final}_box.txt 2>&1
# first call of the round: SETS unset (all variants, coarse prefetch distances); second call: the fine sweep below
# (the first call of the round also had pt_variant 2..7: look-back preload and evict_first loads, removed since)
COARSE='--set "" --set pt_variant=0 --set pt_variant=1,pt_pf_tiles=37 --set pt_variant=1,pt_pf_tiles=148 --set pt_variant=1,pt_pf_tiles=296 --set pt_variant=1,pt_pf_tiles=592'
SETS=${SETS:-$COARSE}
eval timeout 150 python tools/sweep_partition.py --log2n 30 --iters ${ITERS:-3} --out gpurun_out/${TAG:-r2_final}_sweep.json $SETS \
  > gpurun_out/${TAG:-r2_final}_sweep.log 2>&1
cat gpurun_out/${TAG:-r2_final}_sweep.log | cut -c1-220
echo "elapsed $(el)"
TUNE=$(python - <<'PY'
import json
try:
    w = json.load(open("gpurun_out/" + __import__("os").environ.get("TAG", "r2_final") + "_sweep.json"))["winner"] or {}
except Exception:
    w = {}
print(",".join(f"{k}={v}" for k, v in w.items()))
PY
)
echo "candidate: '$TUNE'" | tee gpurun_out/${TAG:-r2_final}_candidate.txt
LSB_TEST_TUNE="$TUNE" timeout 240 python -m pytest tests -m gpu -x -q --ignore=tests/test_multigpu.py > gpurun_out/${TAG:-r2_final}_gpu_tests.log 2>&1
echo "pytest rc=$? (LSB_TEST_TUNE='$TUNE')" | tee -a gpurun_out/${TAG:-r2_final}_gpu_tests.log
tail -4 gpurun_out/${TAG:-r2_final}_gpu_tests.log
echo "elapsed $(el)"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG:-r2_final}_smoke.log 2>&1; tail -1 gpurun_out/${TAG:-r2_final}_smoke.log
TARGS=""
for kv in ${TUNE//,/ }; do TARGS="$TARGS --tune $kv"; done
timeout 200 python bench.py $TARGS > gpurun_out/${TAG:-r2_final}_bench_1gpu.json 2> gpurun_out/${TAG:-r2_final}_bench_1gpu.err
echo "bench rc=$? args='$TARGS'"; tail -c 1500 gpurun_out/${TAG:-r2_final}_bench_1gpu.json; tail -3 gpurun_out/${TAG:-r2_final}_bench_1gpu.err
echo "elapsed $(el)"
if [ -n "$EXTRA" ]; then  # the other 1-GPU points with the candidate: 2^31, radix 8 / 11, 24-bit keys
  for spec in "2p31:--log2n 31" "radix8:--radix 8 --no-alt" "radix11:--radix 11" "mask24:--key-mask 0xFFFFFF"; do
    [ $(el) -lt ${EXTRA_DEADLINE:-170} ] || break
    name=${spec%%:*}; flags=${spec#*:}
    timeout 60 python bench.py $TARGS $flags --steps 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG:-r2_final}_bench_1gpu_$name.json 2>> gpurun_out/${TAG:-r2_final}_bench_1gpu.err
    echo "$name rc=$? elapsed $(el)"
  done
  echo "done elapsed $(el)"; exit 0
fi
if [ -f tools/bin/liblsbsort_prof.so ] && [ $(el) -lt 330 ]; then  # stage clocks per tile: as built, with the L2 prefetch, with the candidate
  ( timeout 40 python tools/prof_stages.py --two-step --log2n 30
    timeout 40 python tools/prof_stages.py --two-step --log2n 30 --tune pt_variant=0 ) > gpurun_out/${TAG:-r2_final}_stage_clocks.txt 2>&1
  cat gpurun_out/${TAG:-r2_final}_stage_clocks.txt | cut -c1-160
  echo "elapsed $(el)"
fi
if [ $(el) -lt 330 ]; then
  timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG:-r2_final}_launches_bench.csv \
    python bench.py $TARGS --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-alt > gpurun_out/${TAG:-r2_final}_ncu_list.log 2>&1
  echo "ncu list rc=$? elapsed $(el)"
fi
if [ $(el) -lt 380 ]; then
  PT=""
  for kv in ${TUNE//,/ }; do PT="$PT --tune $kv"; done
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:"partition_kernel" -c 2 -o gpurun_out/prof_${TAG:-r2_final}_partition_2p30 \
    python tools/prof_sort.py --log2n 30 --iters 1 --no-skip $PT > gpurun_out/${TAG:-r2_final}_ncu_full.log 2>&1
  echo "ncu full rc=$? elapsed $(el)"
fi
echo "done elapsed $(el)"
