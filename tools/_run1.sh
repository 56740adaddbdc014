python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python tools/kbench.py --workload cfg5s --iters 20 | tee gpurun_out/r2z_kbench_mma64_final2.txt
python tools/kbench.py --workload j32k50 --iters 20 | tee -a gpurun_out/r2z_kbench_mma64_final2.txt
