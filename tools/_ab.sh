timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02zl_bench_8gpu.json 2> gpurun_out/r02zl_bench_8gpu.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r02zl_bench_8gpu.json; tail -2 gpurun_out/r02zl_bench_8gpu.err
