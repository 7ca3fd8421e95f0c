set -x
cd $GRAFT_REPO_ROOT
# plain runs first (each must exit 0 without ncu)
python bench.py --steps 100 --warmup 10 > gpurun_out/r01v4_bench_n1.json 2> gpurun_out/r01v4_bench_n1.err || exit 1
python bench.py --workload blend --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r01v4_bench_n1_blend.json 2>/dev/null || exit 1
python bench.py --workload batch --k 100 --steps 20 --warmup 5 > gpurun_out/r01v4_bench_batch.json 2>/dev/null || exit 1
for b in 1 64 128; do python bench.py --workload batch --k 100 --batch $b --steps 30 --warmup 5 > gpurun_out/r01v4_bench_batch_b${b}_k100.json 2>/dev/null || exit 1; done
python bench.py --workload batch --k 20 --batch 1 --steps 100 --warmup 10 > gpurun_out/r01v4_bench_batch_b1_k20.json 2>/dev/null || exit 1
python bench.py --workload binary --steps 200 --warmup 20 > gpurun_out/r01v4_bench_binary.json 2>/dev/null || exit 1
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r01v4_bench_reference.json 2>/dev/null || exit 1
# launch lists
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01v4_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"batch|prep|rerank" -c 24 --csv --log-file gpurun_out/r01v4_launches_batch.csv python bench.py --workload batch --k 100 --steps 2 --warmup 1 > gpurun_out/ncu_l2.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 3 -c 2 -f -o gpurun_out/r01v4_scan python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_gemm_pair -s 3 -c 1 -f -o gpurun_out/r01v4_gemm python bench.py --workload batch --k 100 --steps 2 --warmup 1 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:binary_scan -s 3 -c 1 -f -o gpurun_out/r01v4_binary python bench.py --workload binary --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_rerank_kernel -s 1 -c 1 -f -o gpurun_out/r01v4_rerank python bench.py --workload batch --k 100 --steps 2 --warmup 1 > gpurun_out/ncu_f4.log 2>&1
for r in scan gemm binary rerank; do ncu -i gpurun_out/r01v4_$r.ncu-rep --page raw --csv > gpurun_out/r01v4_${r}_raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -20
