set -x
cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_sort.py -x -q -k "onepass_many or sort_matches_oracle or skewed or each_pass or grouped" 2>&1 | tail -5
timeout 600 python tools/sweep_onepass.py --log2n 30 --iters 2 --set "" --set op_cfg=1,op_t1=232 --set op_cfg=1,op_t1=232,op_hints=15 --set op_cfg=1,op_t1=232,op_nx=4 --set op_cfg=1,op_t1=200 --set op_cfg=1,op_t1=232,op_lead=2,op_nx=4 --set op_t1=200 --set op_hints=15 2>&1 | tail -12
python tools/prof_stages.py --log2n 28 --tune op_cfg=1 --tune op_t1=232
python tools/prof_stages.py --log2n 28
timeout 300 python tools/sweep_onepass.py --log2n 28 --iters 2 --mask 0xFFFFFF --set "" --set op_cfg=1,op_t1=232 2>&1 | tail -3
timeout 300 python tools/sweep_onepass.py --log2n 28 --iters 2 --and-draws 3 --set "" --set op_cfg=1,op_t1=232 2>&1 | tail -3
