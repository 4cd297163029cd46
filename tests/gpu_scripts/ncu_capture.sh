mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-allpairs --skip-module --skip-dp --no-reference-cuda --no-oracle-check"
$B > gpurun_out/r2_bench_prof_plain.json 2> gpurun_out/r2_bench_prof_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_ncu_launch.log 2>&1
tail -2 gpurun_out/r2_ncu_launch.log | cut -c1-200
for s in s1 s2 s3; do
  P="python tests/perf_ops.py --only $s --no-variants --iters 1 --warm 0 --ops vox_fwd_planned,vox_bwd,devox_fwd_planned,devox_bwd_planned"
  $P > gpurun_out/plain_$s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"vox_fwd_cell_kernel|tri_fwd_kernel|tri_bwd_kernel|row_gather_kernel" -c 8 -f -o gpurun_out/prof_r02_$s $P > gpurun_out/ncu_$s.log 2>&1
  tail -1 gpurun_out/ncu_$s.log | cut -c1-150
done
