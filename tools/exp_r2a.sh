set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
# 1. quick correctness first (small), bounded
timeout 300 python -m pytest tests/test_gpu_sort.py -x -q -k "sort_matches_oracle or onepass_many or each_pass or radix_widths or skewed" 2>&1 | tail -15
# 2. timings
timeout 120 python tools/sweep_onepass.py --log2n 28 --iters 2 --set "" --set two_step 2>&1 | tail -5
timeout 400 python tools/sweep_onepass.py --log2n 30 --iters 3 --set "" --set two_step --set op_t1=186 --set op_t1=128 --set op_t1=64 --set op_nx=2 --set op_nx=4 --set op_hints=0 --set op_hints=1 --set op_hints=3 --set op_hints=15 --set op_lead=2,op_nx=4 2>&1 | tail -20
