cd $GRAFT_REPO_ROOT
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29531 tools/prof_mgpu.py --iters 2"
$T --log2n 30 2>&1 | grep "iter 1" | cut -c1-260
$T --log2n 30 --tune ex_ctas=2 2>&1 | grep "iter 1" | cut -c1-260
$T --log2n 30 --tune op_ctas_mgpu=2 --tune ex_ctas=2 2>&1 | grep "iter 1" | cut -c1-260
$T --log2n 30 --tune op_ctas_mgpu=2 --tune ex_threads=512 2>&1 | grep "iter 1" | cut -c1-260
$T --log2n 30 --tune vparts=16 2>&1 | grep "iter 1" | cut -c1-260
$T --log2n 30 --tune vparts=4 2>&1 | grep "iter 1" | cut -c1-260
$T --log2n 30 --two-step --tune ex_threads=512 2>&1 | grep "iter 1" | cut -c1-260
$T --log2n 31 2>&1 | grep "iter 1" | cut -c1-260
$T --log2n 31 --two-step --tune ex_threads=512 2>&1 | grep "iter 1" | cut -c1-260
