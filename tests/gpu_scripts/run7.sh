mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 python -m pytest tests/test_gpu_parity.py -x -q -k "test_trilinear_devoxelize and b2c16n20000r16" > gpurun_out/r2_t7a.log 2>&1; tail -30 gpurun_out/r2_t7a.log | grep -E "py:[0-9]+|Error|passed|failed" | head -20
sed -i 's/os.environ\["CUDA_LAUNCH_BLOCKING"\] = "1"/pass/' tests/gpu_scripts/dbg6.py
python tests/gpu_scripts/dbg6.py 2>&1 | tail -5
