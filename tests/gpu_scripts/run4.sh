set -x
mkdir -p gpurun_out
CMD="python tests/perf_ops.py --only s2 --no-variants --iters 1 --warm 0 --ops devox_bwd_planned"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tri_fwd_kernel|tri_bwd_kernel" -c 2 -f -o gpurun_out/r2_prof4_s2 $CMD > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
