python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for a in "--workload cfg4" "--workload cfg4 --nan 0.01" "--workload cfg4 --nan 0.3" "--workload cfg4 --force-missing" "--workload cfg4 --n 15625" "--workload cfg4 --n 8" "--workload cfg2" "--workload cfg2 --mb 2000" "--workload cfg2 --mb 2000 --nan 0.3" "--workload cfg1" "--workload l88" "--workload cfg4 --mb 100000" "--workload cfg4 --mb 100000 --nan 0.3" "--workload cfg5s" "--workload j32k50" "--workload j64k64"; do
  python tools/kbench.py $a --iters 30
done 2>&1 | tee gpurun_out/r2w_kbench.txt
for w in cfg1 cfg2 cfg3 cfg5; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2w_bench_$w.json 2>/dev/null; done
python bench.py > gpurun_out/r2w_bench_cfg4.json 2> gpurun_out/r2w_bench_cfg4.err
python bench.py --impl reference > gpurun_out/r2w_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_bench_cfg4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_bench_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:elbo_fwdbwd_mma_kernel -s 3 -c 1 -f -o gpurun_out/r2w_mma_kernel python tools/kbench.py --workload cfg4 --iters 2 > gpurun_out/r2w_ncu_full.log 2>&1
compute-sanitizer --tool memcheck python tools/kbench.py --workload cfg4 --n 300 --iters 1 > gpurun_out/r2w_memcheck.txt 2>&1; tail -3 gpurun_out/r2w_memcheck.txt
compute-sanitizer --tool racecheck python tools/kbench.py --workload cfg4 --n 300 --iters 1 > gpurun_out/r2w_racecheck.txt 2>&1; tail -3 gpurun_out/r2w_racecheck.txt
tail -c 1500 gpurun_out/r2w_bench_cfg4.json
