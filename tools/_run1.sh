for rep in 1 2; do
for lib in i12 i02 i11 i01; do
  export CYTOF_B200_LIB=cytofpy_b200/_C/libv_$lib.so
  python tools/kbench.py --workload cfg4 --mb 100000 --iters 30
  python tools/kbench.py --workload cfg4 --mb 100000 --sampler --iters 30
done; done 2>&1 | tee gpurun_out/r2u_kbench_gathered_variants.txt
