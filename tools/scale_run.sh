# The driver's scaling sequence on one multi-GPU box: bench at the given N values (N = 1 is the plain bench).
# Usage (under gpurun --gpus G):  bash tools/scale_run.sh r02 "8 4 2"
TAG=${1:-r02}
NS=${2:-"8 4 2"}
cd $GRAFT_REPO_ROOT
O=gpurun_out
free -g | head -2 > $O/${TAG}_host.txt; nproc >> $O/${TAG}_host.txt; nvidia-smi -L >> $O/${TAG}_host.txt
port=29600
for n in $NS; do
  port=$((port+1))
  timeout 860 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 > $O/${TAG}_bench_n$n.json 2> $O/${TAG}_bench_n$n.err
  echo "N=$n rc=$?"; head -c 400 $O/${TAG}_bench_n$n.json; echo; tail -c 300 $O/${TAG}_bench_n$n.err
done
first=$(echo $NS | cut -d" " -f1)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $first --master-addr 127.0.0.1 --master-port 29650 bench.py --impl reference --gpus $first --steps 20 --warmup 5 > $O/${TAG}_bench_reference_n$first.json 2>/dev/null
echo "ref rc=$?"
