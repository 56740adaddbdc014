python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2w_bench_cfg4_8gpu.json 2> gpurun_out/r2w_bench_cfg4_8gpu.err
tail -c 1200 gpurun_out/r2w_bench_cfg4_8gpu.json
