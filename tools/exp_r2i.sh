set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_sort.py -x -q -k "onepass_many or sort_matches_oracle or skewed or each_pass or grouped" 2>&1 | tail -3
timeout 800 python tools/sweep_onepass.py --log2n 30 --iters 2 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=6,op_hints=15 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=6,op_hints=31 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=5,op_hints=31 --set op_cfg=1,op_t1=232,op_lead=2,op_nx=4,op_hints=31 --set op_cfg=1,op_t1=232,op_lead=4,op_nx=6,op_hints=31 --set op_cfg=1,op_t1=232,op_lead=3,op_nx=8,op_hints=31 2>&1 | tail -6
run() {
  name=$1; shift
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:onepass_kernel -c 1 --csv python tools/prof_sort.py --log2n 28 --iters 1 --no-skip "$@" 2>/dev/null | grep -E "onepass_kernel" | awk -F'","' -v n="$name" '{print n, $(NF-2), $(NF)}' | tr -d '"'
}
run nx6_lead3_h31 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=6 --tune op_lead=3 --tune op_hints=31
run nx5_lead3_h31 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=5 --tune op_lead=3 --tune op_hints=31
run nx4_lead2_h31 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=4 --tune op_lead=2 --tune op_hints=31
run nx6_lead3_h15 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=6 --tune op_lead=3 --tune op_hints=15
