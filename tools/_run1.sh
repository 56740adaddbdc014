python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in default st1 st2; do
  if [ $lib = default ]; then unset CYTOF_B200_LIB; else export CYTOF_B200_LIB=cytofpy_b200/_C/libv_$lib.so; fi
  python tools/kbench.py --workload cfg4
  python tools/kbench.py --workload cfg4 --n 15625
  python tools/kbench.py --workload cfg2 --mb 2000
done 2>&1 | tee gpurun_out/r2o_kbench_variants.txt
unset CYTOF_B200_LIB
python tools/kbench.py --workload cfg4 --n 8 | tee -a gpurun_out/r2o_kbench_variants.txt
python tools/kbench.py --workload cfg1 | tee -a gpurun_out/r2o_kbench_variants.txt
python bench.py --no-cpu-baseline > gpurun_out/r2o_bench_cfg4.json 2> gpurun_out/r2o_bench_cfg4.err; tail -c 600 gpurun_out/r2o_bench_cfg4.json
python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/r2o_bench_cfg2.json 2>/dev/null; tail -c 300 gpurun_out/r2o_bench_cfg2.json
