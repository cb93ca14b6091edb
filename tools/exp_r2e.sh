set -x
cd $GRAFT_REPO_ROOT
timeout 800 python tools/sweep_onepass.py --log2n 30 --iters 2 --set op_lead=2,op_nx=4 --set op_lead=2,op_nx=4,op_t1=200 --set op_lead=2,op_nx=4,op_t1=160 --set op_lead=2,op_nx=5,op_t1=160 --set op_cfg=1,op_t1=232,op_lead=2,op_nx=5 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=5 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=6 --set op_cfg=1,op_t1=232,op_lead=4,op_nx=7 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=6,op_hints=15 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=6,op_hints=3 --set op_cfg=1,op_t1=200,op_lead=3,op_nx=6 2>&1 | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onepass_kernel -c 1 -o gpurun_out/prof_r2_onepass_v2_cfg1_2p27 python tools/prof_sort.py --log2n 27 --iters 1 --no-skip --tune op_cfg=1 --tune op_t1=232 --tune op_lead=3 --tune op_nx=6 > gpurun_out/ncu_r2e.log 2>&1
tail -3 gpurun_out/ncu_r2e.log
