set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refmodel.py -m gpu -x -q -k "voxelize or devox or plan or refmodel or hybrid or heun or inference or pvconv or graph" > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2.log
tail -5 gpurun_out/r2_t2.log
OPS=vox_fwd_planned,devox_fwd_planned,devox_bwd_planned
timeout 300 python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops $OPS > gpurun_out/r2_p2.log 2>&1
for v in pipe0 chunk256 chunk1024 lm32 lm16; do PCFM_B200_LIB=point-cloud-flow-matching_b200/lib/var/libpcfm_$v.so timeout 300 python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops $OPS > gpurun_out/r2_p2_$v.log 2>&1; done
CMD="python tests/perf_ops.py --only s2 --no-variants --iters 1 --warm 0 --ops devox_bwd_planned"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tri_fwd_kernel|tri_bwd_kernel|vox_fwd_cell_kernel" -c 4 -f -o gpurun_out/r2_prof_s2 $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
