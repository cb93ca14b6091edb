set -x
cd $GRAFT_REPO_ROOT
python tools/prof_stages.py --log2n 28
python tools/prof_stages.py --log2n 28 --tune op_t1=64
# ncu full capture of the one-pass kernel at 2^27 (kernel replay restores memory between passes)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onepass_kernel -c 1 -o gpurun_out/prof_r2_onepass_2p27 python tools/prof_sort.py --log2n 27 --iters 1 --no-skip > gpurun_out/ncu_r2b.log 2>&1
tail -5 gpurun_out/ncu_r2b.log
ls -la gpurun_out
