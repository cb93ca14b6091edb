cd $GRAFT_REPO_ROOT
run() {
  name=$1; shift
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:onepass_kernel -c 2 --csv python tools/prof_sort.py --log2n 28 --iters 1 --no-skip "$@" 2>/dev/null | grep -E "onepass_kernel" | awk -F'","' -v n="$name" '{print n, $(NF-2), $(NF)}' | tr -d '"'
}
run cfg0_nx3_lead1
run cfg0_nx4_lead2 --tune op_nx=4 --tune op_lead=2
run cfg1_nx3_lead1 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=3 --tune op_lead=1
run cfg1_nx4_lead2 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=4 --tune op_lead=2
run cfg1_nx5_lead3 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=5 --tune op_lead=3
run cfg1_nx6_lead3 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=6 --tune op_lead=3
run cfg1_nx8_lead3 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=8 --tune op_lead=3
run cfg1_nx6_lead3_h15 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=6 --tune op_lead=3 --tune op_hints=15
run cfg1_nx6_lead3_h0 --tune op_cfg=1 --tune op_t1=232 --tune op_nx=6 --tune op_lead=3 --tune op_hints=0
run cfg1_t116_nx8_lead5 --tune op_cfg=1 --tune op_t1=116 --tune op_nx=8 --tune op_lead=5
