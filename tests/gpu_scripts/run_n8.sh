mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err; head -c 600 gpurun_out/r2_bench_n8.json
