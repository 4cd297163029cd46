mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do echo "dbg=$d"; python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops devox_fwd_planned --opt dbg=$d 2>&1 | cut -c40-110; done > gpurun_out/r2_p5.log 2>&1
echo "tri2=0" >> gpurun_out/r2_p5.log; python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops devox_fwd_planned --opt tri2=0 2>&1 | cut -c40-110 >> gpurun_out/r2_p5.log
cat gpurun_out/r2_p5.log
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "test_trilinear_devoxelize and b2c16n20000r16" > gpurun_out/r2_san5.log 2>&1
grep -E "Invalid|misaligned|at 0x|tri2|========= *at|Error" gpurun_out/r2_san5.log | head -30
