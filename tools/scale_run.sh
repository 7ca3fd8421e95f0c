# The driver's scaling sequence on one 8-GPU box: reference arm + bench at N = 2, 4, 8 (N = 1 is the plain bench).
# Usage (under gpurun --gpus 8):  bash tools/scale_run.sh r02
TAG=${1:-r02}
cd $GRAFT_REPO_ROOT
O=gpurun_out
free -g | head -2 > $O/${TAG}_host.txt; nproc >> $O/${TAG}_host.txt; nvidia-smi -L >> $O/${TAG}_host.txt
port=29600
for n in 8 4 2; do
  port=$((port+1))
  timeout 860 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 > $O/${TAG}_bench_n$n.json 2> $O/${TAG}_bench_n$n.err
  echo "N=$n rc=$?"; head -c 400 $O/${TAG}_bench_n$n.json; echo; tail -c 300 $O/${TAG}_bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > $O/${TAG}_bench_reference_n8.json 2>/dev/null
echo "ref rc=$?"
