set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refmodel.py -m gpu -x -q -k "voxelize or devox or plan or refmodel or hybrid or heun or inference or pvconv or graph" > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
tail -5 gpurun_out/r2_t3.log
OPS=vox_fwd_planned,devox_fwd_planned,devox_bwd_planned,vox_bwd
timeout 300 python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops $OPS > gpurun_out/r2_p3.log 2>&1
cat gpurun_out/r2_p3.log | cut -c40-160
