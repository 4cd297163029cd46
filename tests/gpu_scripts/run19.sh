mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t19.log; tail -4 gpurun_out/r2_t19.log
timeout 1200 python tests/perf_ops.py --sweep --ref --no-variants --iters 5 > gpurun_out/r2_sweep.log 2>&1; tail -2 gpurun_out/r2_sweep.log | cut -c1-300
timeout 900 python tests/perf_ops.py --only pointnet --ref --no-variants --iters 5 > gpurun_out/r2_pointnet.log 2>&1; tail -2 gpurun_out/r2_pointnet.log | cut -c1-300
timeout 600 python tests/perf_ops.py --only s1,s2,s3,c1 --ref --no-variants > gpurun_out/r2_ops_ref.log 2>&1; tail -2 gpurun_out/r2_ops_ref.log | cut -c1-300
ls -la gpurun_out/perf_ops_*.json
