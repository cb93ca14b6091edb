set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu.py -x -q -k "2" 2>&1 | tail -5
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29531 tools/prof_mgpu.py --log2n 30 --iters 2"
$T --tune op_ctas_mgpu=1 2>&1 | grep "iter 1"
$T --tune op_ctas_mgpu=2 2>&1 | grep "iter 1"
$T --tune op_ctas_mgpu=3 2>&1 | grep "iter 1"
$T --tune op_ctas_mgpu=4 2>&1 | grep "iter 1"
$T --tune op_ctas_mgpu=2 --tune op_lead=2 --tune op_nx=4 2>&1 | grep "iter 1"
$T --tune op_ctas_mgpu=3 --tune ex_ctas=2 2>&1 | grep "iter 1"
$T --tune op_ctas_mgpu=2 --tune vparts=16 2>&1 | grep "iter 1"
$T --two-step 2>&1 | grep "iter 1"
