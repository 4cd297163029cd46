mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouping or gather or edge or empty or 3nn or three" > gpurun_out/r2_t24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t24.log
tail -15 gpurun_out/r2_t24.log
for o in 0 1; do echo "== three_nn_bwd=$o"; timeout 600 python tests/perf_ops.py --only pointnet --ops three_nn_bwd,grouping_bwd --opt three_nn_bwd=$o 2>&1 | grep '"op"' | cut -c1-160; done
