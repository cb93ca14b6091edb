set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_sort.py -x -q -k "onepass_many or sort_matches_oracle or skewed or each_pass or grouped or radix_widths or sweep" 2>&1 | tail -5
timeout 800 python tools/sweep_onepass.py --log2n 30 --iters 2 --set "" --set op_lead=2,op_nx=4 --set op_cfg=1,op_t1=232,op_lead=2,op_nx=4 --set op_cfg=1,op_t1=232,op_lead=2,op_nx=4,op_hints=15 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=5,op_hints=15 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=6,op_hints=15 --set op_cfg=1,op_t1=232,op_lead=2,op_nx=3,op_hints=15 --set two_step 2>&1 | tail -9
python tools/prof_stages.py --log2n 28 --tune op_cfg=1 --tune op_t1=232 --tune op_lead=3 --tune op_nx=6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onepass_kernel -c 1 -o gpurun_out/prof_r2_onepass_v3_cfg1_2p27 python tools/prof_sort.py --log2n 27 --iters 1 --no-skip --tune op_cfg=1 --tune op_t1=232 --tune op_lead=3 --tune op_nx=6 --tune op_hints=15 > gpurun_out/ncu_r2g.log 2>&1
tail -2 gpurun_out/ncu_r2g.log
