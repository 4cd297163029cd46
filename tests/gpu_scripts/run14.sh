mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "chamfer or refmodel or hybrid or heun or inference or walk_variants" > gpurun_out/r2_t14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t14.log; tail -4 gpurun_out/r2_t14.log
python tests/perf_ops.py --only chamfer --ref 2>&1 | grep chamfer_fwd | cut -c1-200
python tests/perf_ops.py --only chamfer --opt chamfer_once=0 2>&1 | grep '"chamfer_fwd"' | cut -c1-200
