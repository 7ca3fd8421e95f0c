# Round profile: plain bench runs first (each must exit 0 WITHOUT ncu), then launch lists, then one
# `ncu --set full` capture per hot kernel.  Usage (under gpurun, 1 GPU):  bash tools/profile_round.sh r02
set -x
TAG=${1:-r02}
cd $GRAFT_REPO_ROOT
O=gpurun_out
python bench.py --steps 100 --warmup 10 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err || exit 1
python bench.py --workload batch --k 100 --steps 30 --warmup 5 > $O/${TAG}_bench_batch.json 2>/dev/null || exit 1
python bench.py --workload batch --k 20 --batch 1 --steps 100 --warmup 10 > $O/${TAG}_bench_batch_b1_k20.json 2>/dev/null || exit 1
python bench.py --workload binary --steps 200 --warmup 20 > $O/${TAG}_bench_binary.json 2>/dev/null || exit 1
python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2>/dev/null || exit 1
# launch lists of the same commands (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --sub none --no-cpu-baseline > $O/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"batch|prep|rerank" -c 24 --csv --log-file $O/${TAG}_launches_batch.csv python bench.py --workload batch --k 100 --steps 2 --warmup 1 > $O/ncu_l2.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 3 -c 2 -f -o $O/${TAG}_scan python bench.py --steps 3 --warmup 3 --sub none --no-cpu-baseline > $O/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_gemm_pair -s 3 -c 1 -f -o $O/${TAG}_gemm python bench.py --workload batch --k 100 --steps 2 --warmup 1 > $O/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_gemm_pair -s 3 -c 1 -f -o $O/${TAG}_gemm64 python bench.py --workload batch --k 20 --batch 1 --steps 2 --warmup 1 > $O/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_rerank_kernel -s 1 -c 1 -f -o $O/${TAG}_rerank python bench.py --workload batch --k 100 --steps 2 --warmup 1 > $O/ncu_f4.log 2>&1
for r in scan gemm gemm64 rerank; do ncu -i $O/${TAG}_$r.ncu-rep --page raw --csv > $O/${TAG}_${r}_raw.csv 2>/dev/null; done
rm -f $O/${TAG}_*.ncu-rep.tmp
ls -la $O | tail -20
