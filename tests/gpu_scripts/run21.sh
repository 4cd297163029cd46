mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "devox or walk_variants or bench_shapes or plan" 2>&1 | tail -3
for o in 0 -1 2; do echo "bwd_phases=$o"; python tests/perf_ops.py --only s1,s2,s3 --no-variants --ops devox_bwd_planned --opt bwd_phases=$o 2>&1 | cut -c40-112; done
