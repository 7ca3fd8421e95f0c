cd $GRAFT_REPO_ROOT
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 400 "$@" > $O/var_$name.json 2> $O/var_$name.err; echo "$name rc=$? $(head -c 230 $O/var_$name.json)"; [ -s $O/var_$name.json ] || tail -c 600 $O/var_$name.err; }
run n2_nccl $T --master-port 29701 bench.py --gpus 2 --rows 2000000 --steps 10 --warmup 3 --exchange nccl
run n2_blend $T --master-port 29702 bench.py --gpus 2 --rows 2000000 --steps 10 --warmup 3 --workload blend --sub none
run n2_batch $T --master-port 29703 bench.py --gpus 2 --rows 2000000 --steps 10 --warmup 3 --workload batch
run n2_bf16 $T --master-port 29704 bench.py --gpus 2 --rows 3000000 --steps 10 --warmup 3 --store bf16-primary
run n1_clustered python bench.py --rows 2000000 --steps 5 --warmup 3 --data clustered --no-cpu-baseline
run n1_blend python bench.py --rows 2000000 --steps 5 --warmup 3 --workload blend --no-cpu-baseline
run n1_v2 python bench.py --rows 2000000 --steps 5 --warmup 3 --variant 2 --sub none --no-cpu-baseline
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_exchange.py tests/test_gpu_database.py -x -q -k "not million" 2>&1 | tail -3
