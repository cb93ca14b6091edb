set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_sort.py -x -q -k "sort_matches_oracle or skewed or each_pass or grouped or radix_widths or golden or large" 2>&1 | tail -3
timeout 300 python tools/sweep_onepass.py --log2n 30 --iters 2 --set "" --set two_step 2>&1 | tail -2
timeout 100 python tools/sweep_onepass.py --log2n 28 --iters 2 --mask 0xFFFFFF --set "" 2>&1 | tail -1
timeout 100 python tools/sweep_onepass.py --log2n 28 --iters 2 --radix 14 --set "" 2>&1 | tail -1
timeout 100 python tools/sweep_onepass.py --log2n 28 --iters 2 --radix 15 --set "" 2>&1 | tail -1
