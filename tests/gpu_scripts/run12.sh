mkdir -p gpurun_out
python tests/gpu_scripts/run11_inner.py 2>/dev/null
OPS=vox_fwd_planned,devox_fwd_planned,devox_bwd_planned,vox_bwd
echo "== eager"; python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops $OPS 2>&1 | cut -c40-110
echo "== graph"; python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops $OPS --graph 2>&1 | cut -c40-110
echo "== graph tri2=0"; python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops devox_fwd_planned,devox_bwd_planned --graph --opt tri2=0 2>&1 | cut -c40-110
echo "== graph dbg=3"; python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops devox_fwd_planned --graph --opt dbg=3 2>&1 | cut -c40-110
python -m pytest tests/test_gpu_parity.py -x -q -k "test_trilinear_devoxelize or test_trilinear_walks or plan or voxelize" 2>&1 | tail -3
