mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t13.log; tail -5 gpurun_out/r2_t13.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; tail -3 gpurun_out/r2_bench13.err
