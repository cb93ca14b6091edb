set -x
cd $GRAFT_REPO_ROOT
timeout 100 python -m pytest tests/test_gpu_sort.py -x -q -k "onepass_many or sort_matches_oracle" 2>&1 | tail -3
timeout 600 python tools/sweep_onepass.py --log2n 30 --iters 2 --set "" --set op_persist=48 --set op_persist=80 --set op_cfg=1,op_t1=232 --set op_cfg=1,op_t1=232,op_nx=4 --set op_cfg=1,op_t1=232,op_persist=48 --set op_cfg=1,op_t1=200 --set op_cfg=1,op_t1=232,op_hints=15 --set op_cfg=1,op_t1=232,op_hints=0 --set op_t1=200,op_persist=64 2>&1 | tail -20
LSB_LIBRARY=tools/bin/liblsbsort_prof.so python tools/prof_stages.py --log2n 28 --tune op_cfg=1 --tune op_t1=232
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onepass_kernel -c 1 -o gpurun_out/prof_r2_onepass_cfg1_2p27 python tools/prof_sort.py --log2n 27 --iters 1 --no-skip --tune op_cfg=1 --tune op_t1=232 > gpurun_out/ncu_r2c.log 2>&1
tail -3 gpurun_out/ncu_r2c.log
