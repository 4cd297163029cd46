mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tfinal.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tfinal.log; tail -3 gpurun_out/r2_tfinal.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -1 gpurun_out/r02_bench_n1.err | cut -c1-200
timeout 900 python tests/perf_ops.py --only pointnet --ref --no-variants --iters 5 > /dev/null 2>&1; cp gpurun_out/perf_ops_randn_pointnet.json gpurun_out/r02_perf_pointnet_ops.json
ls gpurun_out/r02_*
timeout 600 python tests/perf_ops.py --only chamfer,emd,allpairs --ref > /dev/null 2>&1; cp gpurun_out/perf_ops_randn_chamfer-emd-allpairs.json gpurun_out/r02_perf_chamfer_emd.json
