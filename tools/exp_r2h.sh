set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_sort.py -x -q -k "onepass_many or sort_matches_oracle or skewed or each_pass or grouped or radix_widths" 2>&1 | tail -3
timeout 800 python tools/sweep_onepass.py --log2n 30 --iters 2 --set op_cfg=1,op_t1=232,op_lead=2,op_nx=4,op_hints=15 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=5,op_hints=15 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=6,op_hints=15 --set op_lead=2,op_nx=4,op_hints=15 --set two_step 2>&1 | tail -6
python tools/prof_stages.py --log2n 28 --tune op_cfg=1 --tune op_t1=232 --tune op_lead=3 --tune op_nx=6
