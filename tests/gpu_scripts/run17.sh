mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "bench_shapes" 2>&1 | grep -E "bench shape|passed|failed|^E " | head
python tests/perf_model.py 2>&1 | grep -E "^\{|Error|error" | cut -c1-400
