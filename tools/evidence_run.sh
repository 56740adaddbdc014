# The GPU-box command behind the final evidence of a round (profiles/r2w_*): full GPU suite, smoke, kernel-alone table,
# bench lines of every BASELINE config + the CPU arm, launch lists of the step, one `ncu --set full` capture of the
# dominant kernel.    gpurun --timeout 1200 -- 'bash tools/evidence_run.sh'
# (2 / 8 GPUs: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N)
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
tail -c 1500 gpurun_out/r2w_bench_cfg4.json
